"""Import the UNMODIFIED reference (`/root/reference`, pure Python) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package, `bench.py` or the
`-m gpu` tests imports this module: `/root/reference` does not exist on the GPU
box.  It is used by `oracle/gen_golden.py` (which writes the committed fixtures
under `tests/golden/`) and by the optional `tests/test_oracle_vs_reference.py`
cross-check that skips itself when `/root/reference` is absent.

Shims, none of which touch the reference's arithmetic:
  1. `ref_shims/matplotlib`  - empty stub (module-top imports only).
  2. `interp2d` -> `RectBivariateSpline(x, y, z.T, kx=3, ky=3, s=0)`; SciPy >= 1.14
     removed interp2d and names this the bug-compatible replacement on regular
     grids (both call FITPACK regrid_smth).  Reference call sites:
     solat_optics/ray_trace_int.py:1417-1425.
  3. `np.complex = complex` (solat_optics/far_field_latrt.py:29).
  4. `ref_shims/pan_mod.py`  - one-panel / zero-error stand-in for holosim-ml.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference in the build container; on the GPU box the byte-for-byte staged copy oracle/_ref (oracle/stage_ref.py)
REF_ROOT = os.environ.get("SOLAT_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/solat_optics") else os.path.join(_HERE, "_ref"))
_REPO = os.path.dirname(_HERE)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "solat_optics"))


class _Interp2dShim:
    def __init__(self, x, y, z, kind="cubic"):
        from scipy.interpolate import RectBivariateSpline

        assert kind == "cubic"
        self._r = RectBivariateSpline(x, y, z.T, kx=3, ky=3, s=0)

    def __call__(self, xnew, ynew):
        return self._r(xnew, ynew).T


def load():
    """Returns a namespace with the reference modules.  Must run in a process that
    has NOT imported this repository's own `solat_optics` drop-in package."""
    import types

    import numpy as np

    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    if "solat_optics" in sys.modules and not str(
        getattr(sys.modules["solat_optics"], "__path__", [""])[0]
    ).startswith(REF_ROOT):
        raise RuntimeError("this process already imported the repo's solat_optics drop-in")
    # the repo root holds a regular package named solat_optics which would shadow the
    # reference's namespace package: drop it (and '') from the path for this process.
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != _REPO]
    sys.path.insert(0, os.path.join(_HERE, "ref_shims"))
    sys.path.insert(0, REF_ROOT)
    if not hasattr(np, "complex"):
        np.complex = complex  # shim 3

    import solat_optics.far_field_latrt as ffl
    import solat_optics.latrt_geo as lgeo
    import solat_optics.ot_geo as ot_geo
    import solat_optics.phase_correct as phase_correct
    import solat_optics.ray_trace_int as rti
    import solat_optics.solat_apert_field as saf

    rti.interp2d = _Interp2dShim  # shim 2
    ns = types.SimpleNamespace(
        ot_geo=ot_geo, rti=rti, lgeo=lgeo, saf=saf, ffl=ffl, phase_correct=phase_correct
    )
    return ns


class PlotOpts:
    """Same role as the class in scripts/holo_sim.py:25-31."""

    plot = 0
    color = "darkorange"
    alpha = 0.2


# ----------------------------------------------------------------------------------
# Worker bodies of the two CLI scripts, driven from a plain function (mpi4py absent).
# ----------------------------------------------------------------------------------
def nf_tele_geo(ref, freq, n_scan, c_light=0.299792458):
    """scripts/holo_sim.py:34-57 (band switch, wavelength) with n_scan made explicit."""
    import numpy as np

    tg = ref.ot_geo.LatGeo()
    if freq <= 120:
        tg.y_ap = -1213 - (6.5 * 10)
        tg.th_fwhp = 40.5 * np.pi / 180
    else:
        tg.y_ap = -1213 - (12.2 * 10)
        tg.th_fwhp = 25 * np.pi / 180
    tg.x_ap = 0
    tg.z_ap = 0
    tg.lambda_ = c_light / freq
    tg.k = 2 * np.pi / tg.lambda_
    tg.n_scan = n_scan
    return tg


def nf_pixel(ref, freq, x, z, n_scan, p_rx=(0, 0, 0), n_scan_rx=None, want_arrays=False):
    """scripts/holo_sim.py:157-185 with ray_trace bound to ray_trace_int."""
    import numpy as np

    tg = nf_tele_geo(ref, freq, n_scan if n_scan_rx is None else n_scan_rx)
    out_rx = ref.rti.rx_to_lyot_model(list(p_rx), tg, PlotOpts).output()
    tg.n_scan = n_scan
    out_so = ref.rti.source_to_lyot_model([x, tg.y_ap, z], tg, PlotOpts).output()
    R_LYOT = 136.54 / 2
    rx_new, so_new = ref.rti.regrid(out_rx, out_so)
    lyot_r = np.sqrt(rx_new[0, :] ** 2 + rx_new[2, :] ** 2)
    beam = (
        so_new[4, :]
        * rx_new[4, :]
        * np.exp(complex(0, 1) * rx_new[3, :] * tg.k / 1e3)
        * np.exp(complex(0, 1) * so_new[3, :] * tg.k / 1e3)
    )
    beam = np.sum(beam[np.where(lyot_r <= R_LYOT)])
    amp = np.abs(beam)
    phi = np.arctan2(np.imag(beam), np.real(beam))
    if want_arrays:
        return amp, phi, out_rx, out_so
    return amp, phi


def ff_tele_geo(ref, freq, offset, n_scan):
    """scripts/ff_sim.py:29-44,64,72-79."""
    import numpy as np

    tg = ref.lgeo.initialize_telescope_geometry()
    tg.rx_x = 0
    tg.rx_z = 0
    tg.rx_y = 0 - offset
    tg.lambda_ = (30.0 / freq) * 0.01
    tg.k = 2 * np.pi / tg.lambda_
    tg.N_scan = n_scan
    tg.x_tow = 0
    tg.y_tow = -7.2
    tg.z_tow = 100e3
    tg.el0 = np.arctan(-tg.y_tow / tg.z_tow)
    return tg


def ff_fan(n_scan):
    import numpy as np

    th = np.linspace(np.pi - 0.00005, np.pi + 0.00005, n_scan)
    ph = np.linspace(0, np.pi, n_scan)
    return th, ph


def ff_pixel(ref, freq, offset, n_scan, az, el, data_x, data_y, e_rec, want_arrays=False):
    """scripts/ff_sim.py:204-221."""
    import numpy as np
    import pan_mod as pm

    tg = ff_tele_geo(ref, freq, offset, n_scan)
    th, ph = ff_fan(n_scan)
    pan2 = pm.panel_model_from_adjuster_offsets(2, np.zeros(1092), 1, 0)
    pan1 = pm.panel_model_from_adjuster_offsets(1, np.zeros(1092), 1, 0)
    apert_rx = ref.saf.ray_mirror_pts(tg, th, ph, el, az)
    f_out = ref.saf.aperature_fields_from_panel_model(pan1, pan2, tg, th, ph, apert_rx, el, az)
    out_new = ref.ffl.project_to_data(f_out, data_x, data_y, tg)
    npts = len(out_new)
    out_beam = np.sum(out_new * e_rec) / npts ** 2
    amp = abs(out_beam)
    phi = np.arctan2(np.imag(out_beam), np.real(out_beam))
    if want_arrays:
        return amp, phi, apert_rx, f_out, out_new
    return amp, phi


def ff_pixel_panels(ref, freq, offset, n_scan, az, el, data_x, data_y, e_rec, pan1, pan2, identify):
    """scripts/ff_sim.py:204-221 with non-trivial panel models: `identify` stands in for pan_mod.identify_panel
    (the reference's own panel arithmetic, saf:587-703,805-852, runs unmodified)."""
    import numpy as np

    tg = ff_tele_geo(ref, freq, offset, n_scan)
    th, ph = ff_fan(n_scan)
    keep = ref.saf.identify_panel
    ref.saf.identify_panel = identify
    try:
        apert_rx = ref.saf.ray_mirror_pts(tg, th, ph, el, az)
        f_out = ref.saf.aperature_fields_from_panel_model(pan1, pan2, tg, th, ph, apert_rx, el, az)
    finally:
        ref.saf.identify_panel = keep
    out_new = ref.ffl.project_to_data(f_out, data_x, data_y, tg)
    npts = len(out_new)
    out_beam = np.sum(out_new * e_rec) / npts ** 2
    return abs(out_beam), np.arctan2(np.imag(out_beam), np.real(out_beam)), apert_rx, f_out
