"""Where does config 4 (25 receiver offsets x near-field map) spend its wall time? (dev probe, GPU)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from solat_optics_b200 import maps

def T(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    print("%-46s %8.3f ms" % (label, best * 1e3), flush=True)
    return out

ax = np.linspace(-60, 60, 5)
pos = [(x, 0.0, z) for x in ax for z in ax]
maps.receiver_offset_sweep(150.0, pos, device=0)
T("receiver_offset_sweep whole", lambda: maps.receiver_offset_sweep(150.0, pos, device=0))
tg = maps.holo_tele_geo(150.0, 100)
tabs = T("  25 x receiver_field", lambda: [maps.receiver_field(tg, tuple(p), None, 100) for p in pos])
opl = T("  np.stack opl", lambda: np.stack([t[0] for t in tabs])); w = np.stack([t[1] for t in tabs])
plan = T("  NearFieldPlan create (25 tables)", lambda: maps.NearFieldPlan(tg, [tg.k], opl, w, 100, 0))
px = maps.holo_pixels()
T("  plan.run", lambda: plan.run(px))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
maps.receiver_offset_sweep(150.0, pos, device=0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
