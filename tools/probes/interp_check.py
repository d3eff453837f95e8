"""Interpolated-cell far-field map (ff_map_interp_kernel) against the exact walk: bit comparison and timing (dev tool, GPU)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
import torch, bench
from solat_optics_b200 import maps
n = int(sys.argv[1]); step = int(sys.argv[2])
tg = maps.ff_tele_geo(140, 1.32)
x5, y5, e5 = bench.e_rec_config5()
side = 2 * bench.NUM + 1
px = np.ascontiguousarray(maps.ff_pixels(tg, bench.RES_ARCMIN, bench.NUM).reshape(side, side, 2)[::step, ::step].reshape(-1, 2))
th, phi = maps.ff_fan(n)
plan = maps.FarFieldPlan(tg, th, phi, x5, y5, e5, device=0)
stream = torch.cuda.current_stream().cuda_stream
d_px = torch.from_numpy(px).cuda()
d_a = torch.empty(len(px), dtype=torch.float64, device="cuda"); d_p = torch.empty_like(d_a)
ms = []
for it in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.run_device(d_px, d_a, d_p, stream); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
np.save(sys.argv[3], np.stack([d_a.cpu().numpy(), d_p.cpu().numpy()]))
print("n %%d px %%d ms %%s status %%d" %% (n, len(px), ["%%.3f" %% m for m in ms], plan.status()))
''' % ROOT
n = sys.argv[1] if len(sys.argv) > 1 else "1000"
step = sys.argv[2] if len(sys.argv) > 2 else "10"
outs = []
for tag, env in (("exact", {"SOLAT_FF_INTERP": "0"}), ("interp", {"SOLAT_FF_INTERP_DEBUG": "1"})):
    e = dict(os.environ); e.update(env)
    f = "/tmp/interp_%s.npy" % tag
    r = subprocess.run([sys.executable, "-c", CHILD, n, step, f], env=e, capture_output=True, text=True)
    print(tag, r.stdout.strip(), r.stderr.strip()[-600:])
    outs.append(f)
import numpy as np
a, b = np.load(outs[0]), np.load(outs[1])
d = np.nonzero((a[0] != b[0]) | (a[1] != b[1]))[0]
side = int(round(np.sqrt(a.shape[1])))
print("differing pixels", len(d), "of", a.shape[1], "first index // side:", sorted(set((d // side).tolist()))[:50], "index % side:", sorted(set((d % side).tolist()))[:50])
print("bit-identical:", np.array_equal(a, b), "max |d amp| %.3e  max |d phase| %.3e" % (np.abs(a[0] - b[0]).max(), np.abs(a[1] - b[1]).max()))
