"""A/B of far-field map kernel variants selected by environment variables at plan creation (dev tool, GPU).

    python tools/ab_ff.py "SOLAT_FF_RPT=1" "SOLAT_FF_RPT=2" ...

Times the device-resident map (CUDA events, median of `reps`) for config 2 (41x41 px x 100^2 rays, bench E_rec) and the
config-5 sub-map (every 10th pixel of 401x401, 1000^2 rays), and checks that all variants agree.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from solat_optics_b200 import maps


def timed(plan, px, reps):
    stream = torch.cuda.current_stream().cuda_stream
    d_px = torch.from_numpy(np.ascontiguousarray(px)).cuda()
    d_a = torch.empty(len(px), dtype=torch.float64, device="cuda")
    d_p = torch.empty_like(d_a)
    for _ in range(3):
        plan.run_device(d_px, d_a, d_p, stream)
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.run_device(d_px, d_a, d_p, stream); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), d_a.cpu().numpy(), d_p.cpu().numpy(), plan.status()


def main():
    variants = sys.argv[1:] or ["SOLAT_FF_RPT=1", "SOLAT_FF_RPT=2"]
    npx, amp, ph = maps.holo_map(140.0, device=0)
    x2, y2, e2 = maps.near_field_to_e_rec(np.c_[npx[:, 0], npx[:, 1], amp[0], ph[0]])
    tg = maps.ff_tele_geo(140, 1.32)
    px2 = maps.ff_pixels(tg, 0.25, 20)
    x5, y5, e5 = bench.e_rec_config5()
    side = 2 * bench.NUM + 1
    px5 = np.ascontiguousarray(maps.ff_pixels(tg, bench.RES_ARCMIN, bench.NUM).reshape(side, side, 2)[::10, ::10].reshape(-1, 2))
    ref = {}
    for v in variants:
        keys = []
        for kv in v.split(","):
            if "=" in kv:
                k, val = kv.split("=", 1)
                os.environ[k] = val
                keys.append(k)
        row = [v]
        for name, args, px, reps in (("cfg2", (maps.ff_fan(100), x2, y2, e2), px2, 20), ("cfg5sub", (maps.ff_fan(1000), x5, y5, e5), px5, 5)):
            (th, phi), x, y, e = args
            plan = maps.FarFieldPlan(tg, th, phi, x, y, e, device=0)
            ms, a, p, st = timed(plan, px, reps)
            plan.close()
            if name not in ref:
                ref[name] = (a, p)
            da = float(np.max(np.abs(a - ref[name][0]) / np.maximum(ref[name][0], 1e-300)))
            dp = float(np.max(np.abs(bench.wrap(p - ref[name][1]))))
            row.append("%s %.4f ms st %d dA %.1e dphi %.1e" % (name, ms, st, da, dp))
        print(" | ".join(row), flush=True)
        for k in keys:
            del os.environ[k]


if __name__ == "__main__":
    main()
