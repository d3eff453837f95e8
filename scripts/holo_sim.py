"""Near-field holography simulation of the LATR optics tube on B200 GPUs.
Output: a .txt file with X, Z, amplitude and phase of the beam (same format as the reference).

To run:   python3 holo_sim.py <frequency_GHz>
          torchrun --nproc-per-node 8 holo_sim.py <frequency_GHz>      (replaces mpiexec -n N)

Drop-in for the reference's scripts/holo_sim.py: same positional argument, same band table
(:40-45), same scan grid RES=51 / SWEEP=100 mm (:47-53), same output path and format (:67-73,
:131-139).  The optional flags default to the reference's hard-coded constants.
"""
import argparse
import time

import numpy as np

import _launch  # noqa: F401  (puts the repository root on sys.path)
from solat_optics_b200 import maps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("freq", type=float, nargs="+", help="frequency [GHz]; several values of one band share a trace")
    ap.add_argument("--res", type=int, default=51)
    ap.add_argument("--sweep", type=float, default=100.0, help="width of the source sweep in 1D [mm]")
    ap.add_argument("--shift", type=float, default=0.0)
    ap.add_argument("--n-scan", type=int, default=100, help="rays per axis of the source fan (tele_geo.n_scan)")
    ap.add_argument("--n-scan-rx", type=int, default=None, help="rays per axis of the receiver fan (default: --n-scan)")
    ap.add_argument("--rx", type=float, nargs=3, default=(0.0, 0.0, 0.0), help="receiver feed position [mm]")
    ap.add_argument("--out", default=None)
    ap.add_argument("--strict", action="store_true", help="raise where the reference's brentq would")
    a = ap.parse_args()

    rank, world = _launch.init_distributed()
    t0 = time.perf_counter()
    res = maps.holo_map(a.freq, a.res, a.sweep, a.shift, a.n_scan, a.n_scan_rx, tuple(a.rx), strict=a.strict)
    if res is not None:
        pixels, amp, phase = res
        for i, f in enumerate(a.freq):
            tg = maps.holo_tele_geo(f, a.n_scan)
            name = a.out if (a.out and len(a.freq) == 1) else (
                "data/source_test_" + str(round(-tg.y_ap / 1e3, 2)) + "m_" + str(int(f)) + "GHz.txt")
            maps.save_map_txt(name, pixels[:, 0], pixels[:, 1], amp[i], phase[i])
            print("wrote %s  (%d px x %d rays, %d GPU(s), %.2f s)" % (
                name, len(pixels), a.n_scan ** 2, world, time.perf_counter() - t0))
    _launch.finish_distributed()


if __name__ == "__main__":
    main()
