"""Device time of the config-5 sub-map (every 10th pixel, 1000^2 rays) under the current environment (dev tool, GPU)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
from solat_optics_b200 import maps
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
tg = maps.ff_tele_geo(140, 1.32)
x5, y5, e5 = bench.e_rec_config5()
side = 2 * bench.NUM + 1
px = np.ascontiguousarray(maps.ff_pixels(tg, bench.RES_ARCMIN, bench.NUM).reshape(side, side, 2)[::10, ::10].reshape(-1, 2))
th, phi = maps.ff_fan(n)
plan = maps.FarFieldPlan(tg, th, phi, x5, y5, e5, device=0)
stream = torch.cuda.current_stream().cuda_stream
d_px = torch.from_numpy(px).cuda()
d_a = torch.empty(len(px), dtype=torch.float64, device="cuda"); d_p = torch.empty_like(d_a)
ms = []
for it in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.run_device(d_px, d_a, d_p, stream); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
print("%s n %d: median %.3f ms  sumA %.12g sumP %.12g status %d" % (os.environ.get("TAG", ""), n, float(np.median(ms[2:])), float(d_a.sum()), float(d_p.abs().sum()), plan.status()))
