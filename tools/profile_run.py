"""Short single-GPU driver for ncu: a few launches of the fused map kernels at the bench sizes."""
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solat_optics_b200 import maps

which = sys.argv[1] if len(sys.argv) > 1 else "both"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tgn = maps.holo_tele_geo(140.0, 100)
rx_opl, rx_w = maps.receiver_field(tgn)
nplan = maps.NearFieldPlan(tgn, [tgn.k], rx_opl, rx_w, device=0)
px = maps.holo_pixels()
for _ in range(reps if which in ("nf", "both") else 1):
    amp, ph, st = nplan.run(px)
print("nf status", st, "active rays", nplan.active_rays, "amp[centre]", amp[0, 1300])
if which in ("ff", "both"):
    x, y, e = maps.near_field_to_e_rec(np.c_[px[:, 0], px[:, 1], amp[0], ph[0]])
    tg = maps.ff_tele_geo(140, 1.32)
    th, phi = maps.ff_fan(100)
    plan = maps.FarFieldPlan(tg, th, phi, x, y, e, device=0)
    fpx = maps.ff_pixels(tg, 0.25, 20)
    for _ in range(reps):
        a, p, st = plan.run(fpx)
    print("ff status", st, "amp[centre]", a[840])
