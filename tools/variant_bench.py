"""Time and accuracy-check one build variant of libsolat_b200.so (dev tool; run on the GPU box).
usage: SOLAT_B200_LIB=/path/to/variant.so python tools/variant_bench.py <label>"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from solat_optics_b200 import maps, _lib
from oracle import oracle_lib as orc

label = sys.argv[1] if len(sys.argv) > 1 else "?"
reps = 20
def wrap(a): return (a + np.pi) % (2 * np.pi) - np.pi

def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]

stream = torch.cuda.current_stream().cuda_stream
# near field config 1
tgn = maps.holo_tele_geo(140.0, 100)
rx_opl, rx_w = maps.receiver_field(tgn)
nplan = maps.NearFieldPlan(tgn, [tgn.k], rx_opl, rx_w, device=0)
px = maps.holo_pixels()
d_px = torch.from_numpy(px).cuda(); d_a = torch.empty(len(px), dtype=torch.float64, device="cuda"); d_p = torch.empty_like(d_a)
nf_med, nf_min = timed(lambda: nplan.run_device(d_px, d_a, d_p, stream))
amp, ph, st_n = nplan.run(px)
samp = np.arange(0, len(px), 37)
o_rx = orc.nf_rx_trace([0, 0, 0], tgn.th_fwhp, 100)
oa, op = orc.nf_map(px[samp], tgn.y_ap, 100, o_rx[3], o_rx[4] * orc.lyot_mask(100), [tgn.k])
nf_dp = np.max(np.abs(amp[0, samp] ** 2 - oa[0] ** 2) / oa[0] ** 2); nf_df = np.max(np.abs(wrap(ph[0, samp] - op[0])))
# far field config 2
x, y, e = maps.near_field_to_e_rec(np.c_[px[:, 0], px[:, 1], amp[0], ph[0]])
tg = maps.ff_tele_geo(140, 1.32)
th, phi = maps.ff_fan(100)
plan = maps.FarFieldPlan(tg, th, phi, x, y, e, device=0)
fpx = maps.ff_pixels(tg, 0.25, 20)
d_fp = torch.from_numpy(fpx).cuda(); d_fa = torch.empty(len(fpx), dtype=torch.float64, device="cuda"); d_fph = torch.empty_like(d_fa)
ff_med, ff_min = timed(lambda: plan.run_device(d_fp, d_fa, d_fph, stream))
fa, fph, st_f = plan.run(fpx)
fs = np.arange(0, len(fpx), 23)
ofa, ofp = orc.ff_map(orc.ff_geo_vec(rx_y_m=-1.32), fpx[fs], th, phi, x, y, e, tg.k)
ff_dp = np.max(np.abs(fa[fs] ** 2 - ofa ** 2) / ofa ** 2); ff_df = np.max(np.abs(wrap(fph[fs] - ofp)))
print("VARIANT %-22s NF %.3f ms (min %.3f) dP %.1e dphi %.1e st %d | FF %.3f ms (min %.3f) dP %.1e dphi %.1e st %d"
      % (label, nf_med, nf_min, nf_dp, nf_df, st_n, ff_med, ff_min, ff_dp, ff_df, st_f))
if label == "math":
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(0.5, 2.0, 20000), 10 ** rng.uniform(-4, 9, 20000)])
    o = _lib.selftest_math(xs)
    for i, nm in enumerate(["rcp0", "rcp1", "rcp2", "rcp3"]):
        print("  %s max rel err %.3e" % (nm, np.max(np.abs(o[i] * xs - 1))))
    for i, nm in [(4, "rsqrt0"), (5, "rsqrt1")]:
        print("  %s max rel err %.3e" % (nm, np.max(np.abs(o[i] * np.sqrt(xs) - 1))))
    big = rng.uniform(2.8e8, 3.0e8, 20000)
    o = _lib.selftest_math(big)
    print("  cis_large max abs err cos %.2e sin %.2e" % (np.max(np.abs(o[6] - np.cos(big))), np.max(np.abs(o[7] - np.sin(big)))))
if label == "overhead":
    # per-pixel fixed cost of the far-field kernel: the same 1681 pixels with fans of 2x2 ... 100x100 rays
    for n in (2, 10, 30, 50, 70, 100):
        t, p_ = maps.ff_fan(n)
        pl = maps.FarFieldPlan(tg, t, p_, x, y, e, device=0)
        med, mn = timed(lambda: pl.run_device(d_fp, d_fa, d_fph, stream))
        print("  fan %3d x %3d: %.3f ms (min %.3f)" % (n, n, med, mn))
