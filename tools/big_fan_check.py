"""Config-5-style check (dev tool, GPU): 10^6 rays per pixel on a 101x101 closed-form E_rec grid.
Times a few hundred pixels and checks a handful against the oracle (the oracle needs ~2.5 s per pixel)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from solat_optics_b200 import maps
from oracle import oracle_lib as orc

n_scan = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
npix_time = int(sys.argv[2]) if len(sys.argv) > 2 else 444
ax = np.linspace(-5, 5, 101)
y, x = np.meshgrid(ax, ax, indexing="ij")
e = np.exp(-(x ** 2 + y ** 2) / (2 * 1.5 ** 2)) * np.exp(1j * 0.3 * (x ** 2 + y ** 2))
tg = maps.ff_tele_geo(140, 1.32)
th, ph = maps.ff_fan(n_scan)
plan = maps.FarFieldPlan(tg, th, ph, x, y, e, device=0)
px = maps.ff_pixels(tg, 0.05, 200)                      # 401 x 401 grid
sel = px[np.linspace(0, len(px) - 1, npix_time).astype(int)]
d_px = torch.from_numpy(sel).cuda(); d_a = torch.empty(len(sel), dtype=torch.float64, device="cuda"); d_p = torch.empty_like(d_a)
s = torch.cuda.current_stream().cuda_stream
plan.run_device(d_px, d_a, d_p, s); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.run_device(d_px, d_a, d_p, s); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
rays = len(sel) * n_scan ** 2
print("big fan: %d px x %d rays in %.2f ms -> %.3e pixel*rays/s; full 401x401 map extrapolates to %.2f s on 1 GPU; status %d"
      % (len(sel), n_scan ** 2, ms, rays / ms * 1e3, 160801 * n_scan ** 2 / (rays / ms * 1e3), plan.status()))
chk = sel[[0, len(sel) // 2, len(sel) - 1]]
t0 = time.time()
oa, op = orc.ff_map(orc.ff_geo_vec(rx_y_m=-1.32), chk, th, ph, x, y, e, tg.k)
a = d_a.cpu().numpy()[[0, len(sel) // 2, len(sel) - 1]]; p = d_p.cpu().numpy()[[0, len(sel) // 2, len(sel) - 1]]
wrap = lambda v: (v + np.pi) % (2 * np.pi) - np.pi
print("oracle check (%.1f s): dP/P %s dphi %s" % (time.time() - t0, np.abs(a ** 2 - oa ** 2) / oa ** 2, np.abs(wrap(p - op))))
