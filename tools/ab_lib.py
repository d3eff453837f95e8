"""A/B of prebuilt library variants (dev tool, GPU):  python tools/ab_lib.py gpurun_variants/a.so gpurun_variants/b.so ...
Each variant runs in its own process (SOLAT_B200_LIB selects the library); prints the device-resident map times
(median, CUDA events) of config 1 (near field), config 2 and the config-5 sub-map, and a checksum of the results."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
import torch, bench
from solat_optics_b200 import maps
def timed(plan, px, n, reps):
    stream = torch.cuda.current_stream().cuda_stream
    d_px = torch.from_numpy(np.ascontiguousarray(px)).cuda()
    d_a = torch.empty(n, dtype=torch.float64, device="cuda"); d_p = torch.empty_like(d_a)
    for _ in range(3): plan.run_device(d_px, d_a, d_p, stream)
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.run_device(d_px, d_a, d_p, stream); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(d_a.sum()), float(d_p.abs().sum()), plan.status()
out = []
tgn = maps.holo_tele_geo(140.0, 100)
opl, w = maps.receiver_field(tgn)
nplan = maps.NearFieldPlan(tgn, [tgn.k], opl, w, device=0)
npx = maps.holo_pixels()
out.append(("cfg1",) + timed(nplan, npx, len(npx), 30))
amp, ph, _ = nplan.run(npx)
x2, y2, e2 = maps.near_field_to_e_rec(np.c_[npx[:, 0], npx[:, 1], amp[0], ph[0]])
tg = maps.ff_tele_geo(140, 1.32)
px2 = maps.ff_pixels(tg, 0.25, 20)
th, phi = maps.ff_fan(100)
out.append(("cfg2",) + timed(maps.FarFieldPlan(tg, th, phi, x2, y2, e2, device=0), px2, len(px2), 30))
x5, y5, e5 = bench.e_rec_config5()
side = 2 * bench.NUM + 1
px5 = np.ascontiguousarray(maps.ff_pixels(tg, bench.RES_ARCMIN, bench.NUM).reshape(side, side, 2)[::10, ::10].reshape(-1, 2))
th, phi = maps.ff_fan(1000)
out.append(("cfg5sub",) + timed(maps.FarFieldPlan(tg, th, phi, x5, y5, e5, device=0), px5, len(px5), 7))
print(" | ".join("%%s %%.4f ms (sumA %%.12g sumP %%.12g st %%d)" %% o for o in out))
''' % ROOT

for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["SOLAT_B200_LIB"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print("%-34s %s" % (os.path.basename(lib), r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
