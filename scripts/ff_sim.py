"""Far-field beam simulation of the LAT (two mirrors) on B200 GPUs.

To run:   python3 ff_sim.py <freq_GHz> <res_arcmin> <num> <offset_m>
          torchrun --nproc-per-node 8 ff_sim.py 150 1 10 1.33        (replaces mpiexec -n 10)

Drop-in for the reference's scripts/ff_sim.py: same four positional arguments (:23-26), same az/el
grid (:94,121-128), source 100 km away (:72-79), E_rec read from a near-field .txt (:46-59,81-88),
same output format (:181-189).  The reference's hard-coded absolute paths (:46-49) become
--nf-input / --out with repository-relative defaults.
"""
import argparse
import time

import numpy as np

import _launch  # noqa: F401
from solat_optics_b200 import maps


def panel_models(a):
    """ff_sim.py:61-70: panel models from the external pan_mod, if the user has it on PYTHONPATH."""
    if a.panels == "off":
        return None
    try:
        import pan_mod as pm
    except ImportError:
        if a.adj_offsets:
            raise SystemExit("--adj-offsets needs holosim-ml's pan_mod on PYTHONPATH")
        return None
    adj = np.load(a.adj_offsets) if a.adj_offsets else np.zeros((2, 1092))
    save = 0
    return pm.panel_model_from_adjuster_offsets(1, adj[0], 1, save), pm.panel_model_from_adjuster_offsets(2, adj[1], 1, save)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("freq", type=int)
    ap.add_argument("res", type=float, help="pixel size [arcmin]")
    ap.add_argument("num", type=int, help="pixels each side of the centre (map is (2 num + 1)^2)")
    ap.add_argument("offset", type=float, help="focal-plane offset [m]")
    ap.add_argument("--nf-input", default=None, help="near-field map (.txt from holo_sim.py)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--n-scan", type=int, default=100)
    ap.add_argument("--strict", action="store_true",
                    help="raise where the reference's brentq would (a root outside its bracket).  Rays the projection "
                         "would skip anyway are not traced and report no status (fans of 400^2 rays and more: only the "
                         "lattice rays and the winner of every cell are): set SOLAT_FF_PREFILTER=0 to trace and check "
                         "every accepted ray as the reference does (saf:706)")
    ap.add_argument("--panels", choices=["auto", "off"], default="auto",
                    help="auto: when holosim-ml's pan_mod is importable, build the panel models as ff_sim.py:61-70 does "
                         "(identify_panel on the host, panel errors in the kernel); off: every ray on an error-free panel")
    ap.add_argument("--adj-offsets", default=None,
                    help=".npy with the adjuster offsets (2, 1092) [M1, M2] fed to pan_mod.panel_model_from_adjuster_offsets "
                         "(default: zeros, ff_sim.py:61-62)")
    a = ap.parse_args()

    # default input: the file `holo_sim.py <freq>` of this package writes (the reference's own default, ff_sim.py:36-49,
    # names a 1.25 m / 1.32 m file under the author's home directory that no script in the repository produces)
    y_ap = maps.holo_tele_geo(float(a.freq)).y_ap
    nf_path = a.nf_input or "data/source_test_%sm_%dGHz.txt" % (round(-y_ap / 1e3, 2), a.freq)
    out_path = a.out or "output_files/ff_%sm_%dGHz.txt" % (a.offset, a.freq)

    rank, world = _launch.init_distributed()
    t0 = time.perf_counter()
    data_x, data_y, e_rec = maps.near_field_to_e_rec(np.loadtxt(nf_path))
    res = maps.ff_map(a.freq, a.res, a.num, a.offset, data_x, data_y, e_rec, a.n_scan, strict=a.strict,
                      panel_models=panel_models(a))
    if res is not None:
        pixels, amp, phase = res
        maps.save_map_txt(out_path, pixels[:, 0], pixels[:, 1], amp, phase)
        print("wrote %s  (%d px x %d rays, %d GPU(s), %.2f s)" % (
            out_path, len(pixels), a.n_scan ** 2, world, time.perf_counter() - t0))
    _launch.finish_distributed()


if __name__ == "__main__":
    main()
