"""A few 25-table near-field maps (BASELINE config 4 geometry) for ncu (dev probe, GPU)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from solat_optics_b200 import maps

tg = maps.holo_tele_geo(150.0, 100)
ax = np.linspace(-60, 60, 5)
pos = [(x, 0.0, z) for x in ax for z in ax]
opl, w = maps.receiver_fields(tg, pos, 100, 0)
plan = maps.NearFieldPlan(tg, [tg.k], opl, w, 100, 0)
px = maps.holo_pixels()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    a, p, st = plan.run(px)
print("status", st, "amp[12, 0, 1300]", a[12, 0, 1300])
