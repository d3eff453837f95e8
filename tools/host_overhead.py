"""Where does the wall time of a whole map through the public API go? (dev tool, GPU)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solat_optics_b200 import maps, _lib

def T(label, fn, reps=5):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    print("%-46s %8.3f ms" % (label, best * 1e3))
    return out

maps.holo_map(140.0, device=0)
T("holo_map(140) whole call", lambda: maps.holo_map(140.0, device=0))
tg = maps.holo_tele_geo(140.0, 100)
T("  holo_tele_geo", lambda: maps.holo_tele_geo(140.0, 100))
rx = T("  receiver_field (rx trace + regrid + mask)", lambda: maps.receiver_field(tg))
plan = T("  NearFieldPlan create", lambda: maps.NearFieldPlan(tg, [tg.k], rx[0], rx[1], 100, 0))
px = T("  holo_pixels", lambda: maps.holo_pixels())
T("  plan.run (host buffers)", lambda: plan.run(px))
T("  plan.close + recreate", lambda: (plan.close(), maps.NearFieldPlan(tg, [tg.k], rx[0], rx[1], 100, 0))[1])
px, amp, ph = maps.holo_map(140.0, device=0)
x, y, e = T("near_field_to_e_rec (incl. unwrap)", lambda: maps.near_field_to_e_rec(np.c_[px[:, 0], px[:, 1], amp[0], ph[0]]))
T("ff_map whole call", lambda: maps.ff_map(140, 0.25, 20, 1.32, x, y, e, device=0))
tgf = maps.ff_tele_geo(140, 1.32)
th, phi = maps.ff_fan(100)
fplan = T("  FarFieldPlan create", lambda: maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0))
fpx = maps.ff_pixels(tgf, 0.25, 20)
T("  plan.run (host buffers)", lambda: fplan.run(fpx))
