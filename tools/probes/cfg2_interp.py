"""Config 2 (41x41 px x 100^2 rays on the build's near-field E_rec) under the current environment: time + checksum (dev tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from solat_optics_b200 import maps
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
tgn = maps.holo_tele_geo(140.0, 100)
opl, w = maps.receiver_field(tgn)
nplan = maps.NearFieldPlan(tgn, [tgn.k], opl, w, device=0)
npx = maps.holo_pixels()
amp, ph, _ = nplan.run(npx)
x2, y2, e2 = maps.near_field_to_e_rec(np.c_[npx[:, 0], npx[:, 1], amp[0], ph[0]])
tg = maps.ff_tele_geo(140, 1.32)
px = maps.ff_pixels(tg, 0.25, 20)
th, phi = maps.ff_fan(n)
plan = maps.FarFieldPlan(tg, th, phi, x2, y2, e2, device=0)
stream = torch.cuda.current_stream().cuda_stream
d_px = torch.from_numpy(np.ascontiguousarray(px)).cuda()
d_a = torch.empty(len(px), dtype=torch.float64, device="cuda"); d_p = torch.empty_like(d_a)
ms = []
for it in range(30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.run_device(d_px, d_a, d_p, stream); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
np.save(os.environ.get("OUT", "/tmp/cfg2.npy"), np.stack([d_a.cpu().numpy(), d_p.cpu().numpy()]))
print("%s n %d: median %.4f ms  sumA %.12g sumP %.12g status %d" % (os.environ.get("TAG", ""), n, float(np.median(ms[3:])), float(d_a.sum()), float(d_p.abs().sum()), plan.status()))
