"""Short single-GPU driver for ncu: a few launches of the timed kernels of one workload (device-resident path).

    python tools/profile_run.py cfg5 [reps]     # BASELINE config 5: every 10th pixel of the 401 x 401 map, 1000^2 rays
    python tools/profile_run.py cfg2 [reps]     # config 2: 41 x 41 px x 100^2 rays on the build's own near-field E_rec
    python tools/profile_run.py cfg1 [reps]     # config 1: near-field 51 x 51 x 100^2 rays
    python tools/profile_run.py cfg4 [reps]     # config 4: 25 receiver tables, trace once + contraction

The per-pixel*ray figures of a capture go into profiles/r02_kernels.json (tools/ncu_to_profile.py); bench.py scales
them to the full map.  PIXEL_RAYS below is what one launch of the workload processes.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from solat_optics_b200 import maps

PIXEL_RAYS = {"cfg5": 41 * 41 * 10 ** 6, "cfg2": 1681 * 10 ** 4, "cfg1": 2601 * 10 ** 4, "cfg4": 25 * 2601 * 10 ** 4}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    stream = torch.cuda.current_stream().cuda_stream
    if which in ("cfg1", "cfg4"):
        freq = 140.0 if which == "cfg1" else 150.0
        tg = maps.holo_tele_geo(freq, 100)
        if which == "cfg1":
            opl, w = maps.receiver_field(tg)
        else:
            ax = np.linspace(-60, 60, 5)
            opl, w = maps.receiver_fields(tg, [(a, 0.0, b) for a in ax for b in ax], 100, 0)
        plan = maps.near_field_plan(tg, [tg.k], opl, w, 100, 0)
        px = maps.holo_pixels()
        n = plan.channels * len(px)
    else:
        if which == "cfg2":
            npx, amp, ph = maps.holo_map(140.0, device=0)
            x, y, e = maps.near_field_to_e_rec(np.c_[npx[:, 0], npx[:, 1], amp[0], ph[0]])
            tg = maps.ff_tele_geo(140, 1.32)
            th, phi = maps.ff_fan(100)
            px = maps.ff_pixels(tg, 0.25, 20)
        else:
            x, y, e = bench.e_rec_config5()
            tg = maps.ff_tele_geo(bench.FREQ, bench.OFFSET)
            th, phi = maps.ff_fan(bench.N_SCAN)
            side = 2 * bench.NUM + 1
            full = maps.ff_pixels(tg, bench.RES_ARCMIN, bench.NUM).reshape(side, side, 2)
            px = np.ascontiguousarray(full[::10, ::10].reshape(-1, 2))
        plan = maps.far_field_plan(tg, th, phi, x, y, e, 0)
        n = len(px)
    d_px = torch.from_numpy(np.ascontiguousarray(px)).cuda()
    d_a = torch.empty(n, dtype=torch.float64, device="cuda")
    d_p = torch.empty_like(d_a)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        plan.run_device(d_px, d_a, d_p, stream)
        ev[i + 1].record()
    torch.cuda.synchronize()
    print(which, "pixels", len(px), "pixel_rays", PIXEL_RAYS[which], "status", plan.status(),
          "ms", ["%.4f" % ev[i].elapsed_time(ev[i + 1]) for i in range(reps)], "amp[mid]", float(d_a[n // 2]))


if __name__ == "__main__":
    main()
