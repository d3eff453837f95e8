"""Near-field map time against the number of accumulators a launch carries (dev probe, GPU).
SOLAT_NF_MAX_SUMS / SOLAT_NF_QUADS / SOLAT_NF_PAIRS are read per plan / per launch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from solat_optics_b200 import maps

tg = maps.holo_tele_geo(150.0, 100)
ax = np.linspace(-60, 60, 5)
pos = [(x, 0.0, z) for x in ax for z in ax]
opl, w = maps.receiver_fields(tg, pos, 100, 0)
px = maps.holo_pixels()
k3 = [tg.k, maps.holo_tele_geo(220.0, 100).k, maps.holo_tele_geo(280.0, 100).k]


def best(fn, reps=5):
    b = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); b = min(b, time.perf_counter() - t0)
    return b * 1e3


for mode in ("quads", "pairs", "none"):
    os.environ.pop("SOLAT_NF_QUADS", None); os.environ.pop("SOLAT_NF_PAIRS", None)
    if mode == "pairs":
        os.environ["SOLAT_NF_QUADS"] = "0"
    if mode == "none":
        os.environ["SOLAT_NF_PAIRS"] = "0"
    p25 = maps.NearFieldPlan(tg, [tg.k], opl, w, 100, 0)
    p3 = maps.NearFieldPlan(tg, k3, opl[12], w[12], 100, 0)
    for sums in (4, 8, 16, 32):
        os.environ["SOLAT_NF_MAX_SUMS"] = str(sums)
        p25.run(px); p3.run(px)
        print("%-5s max_sums %2d: 25 tables %.3f ms, 3 wavenumbers %.3f ms" % (mode, sums, best(lambda: p25.run(px)), best(lambda: p3.run(px))), flush=True)
    p25.close(); p3.close()
