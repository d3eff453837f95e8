"""Small end-to-end pass over every kernel for compute-sanitizer (dev tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solat_optics_b200 import maps, ray_trace_int as rti, solat_apert_field as saf, far_field_latrt as ffl, _lib

class PlotOpts:
    plot = 0

tg = maps.holo_tele_geo(140.0, 12)
print(rti.rx_to_lyot_model([0, 0, 0], tg, PlotOpts).output().shape, rti.source_to_lyot_model([5, tg.y_ap, -3], tg, PlotOpts).output().shape,
      rti.aperature_fields([0, 0, 0], tg, PlotOpts).output().shape)
px, amp, ph = maps.holo_map([150.0, 220.0], res=4, n_scan=17, n_scan_rx=9)
print("nf", amp.shape, float(amp[0, 0]))
px, a, p = maps.receiver_offset_sweep(150.0, [(0, 0, 0), (30, 0, -20), (-40, 0, 10)], res=3, n_scan=13)
print("sweep", a.shape)
tgf = maps.ff_tele_geo(140, 1.32)
th, phi = maps.ff_fan(21)
m = saf.ray_mirror_pts(tgf, th, phi, 7.2e-5, 1e-4)
pm = saf.panel_backend().panel_model_from_adjuster_offsets(1, None, 1, 0)
f = saf.aperature_fields_from_panel_model(pm, pm, tgf, th, phi, m, 7.2e-5, 1e-4)
ax = np.linspace(-30, 30, 33)
y, x = np.meshgrid(ax, ax, indexing="ij")
e = np.exp(-(x ** 2 + y ** 2) / 200) * np.exp(0.01j * x * y)
print("project", np.count_nonzero(ffl.project_to_data(f, x, y, tgf)))
px, a, p = maps.ff_map(140, 0.25, 2, 1.32, x, y, e, n_scan=37)
print("ff", a.shape, float(a[12]))
print("selftest", _lib.selftest_math(np.linspace(0.5, 3e8, 100)).shape, "peak", _lib.context(0).fp64_peak_tflops(1))
