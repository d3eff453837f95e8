"""Near-field ray tracers: drop-in for the reference's `solat_optics/ray_trace_int.py`.

Classes keep the reference constructor `X(p, tele_geo, plotOpts)`, the attributes
(.tele_geo, .theta, .phi, .th_fwhp, .y_ap, .n_pts, .out), `.trace_rays(ii)` and `.output()`
with the reference's row layout.  The per-ray Python/brentq loop of the reference
(rti:878-885, 1375-1382, 567-576) is replaced by one CUDA launch (solat_nf_*_trace).

`strict` (module attribute, default False): when True, `.output()` raises ValueError where the
reference's brentq would (a ray's root outside the reference bracket).
"""
import ctypes

import numpy as np

from . import _lib
from . import ot_geo  # noqa: F401  (the reference re-exports ot_geo names from this module)
from .ot_geo import *  # noqa: F401,F403

strict = False


def snell_vec(n_1, n_2, n_surf, s_1):
    """Vector Snell's law (reference rti:20-30): s_1 points from the source to the surface,
    n_surf is the surface normal.  Host utility; the kernels use the cross-product-free form."""
    n_surf = np.asarray(n_surf, dtype=float)
    s_1 = np.asarray(s_1, dtype=float)
    mu = n_1 / n_2
    tangential = np.cross(n_surf, np.cross(-n_surf, s_1))
    sin2 = np.dot(np.cross(n_surf, s_1), np.cross(n_surf, s_1))
    return mu * tangential - n_surf * np.sqrt(1 - mu ** 2 * sin2)


class _LyotTracer:
    _half_width = None      # launch-fan half widths about the centres below [rad]
    _theta0 = None
    _phi0 = None
    _rows = None
    _entry = None

    def __init__(self, p, tele_geo, plotOpts):
        n_linear = tele_geo.n_scan
        self._theta_axis = np.linspace(self._theta0 - self._half_width, self._theta0 + self._half_width, n_linear)
        self._phi_axis = np.linspace(self._phi0 - self._half_width, self._phi0 + self._half_width, n_linear)
        theta, phi = np.meshgrid(self._theta_axis, self._phi_axis)
        self.tele_geo = tele_geo
        self.theta = np.ravel(theta)
        self.phi = np.ravel(phi)
        self.th_fwhp = tele_geo.th_fwhp
        self.y_ap = tele_geo.y_ap
        self.n_pts = len(self.theta)
        self._p = p
        self.out = np.zeros((17, self.n_pts))
        self.status = None
        self._cache = None

    def _run(self, theta, nth, phi, nph):
        lib = _lib.load()
        ctx = _lib.context()
        geom = _lib.nf_geom(self.tele_geo)
        p = _lib.as_f64([self._p[0], self._p[1], self._p[2]])
        theta = _lib.as_f64(theta)
        phi = _lib.as_f64(phi)
        n = nth if nph == 0 else nth * nph
        out = np.empty((self._rows, n), dtype=np.float64)
        status = np.empty(n, dtype=np.int32)
        fn = getattr(lib, self._entry)
        _lib.check(fn(ctx.handle, ctypes.byref(geom), _lib.dptr(p), _lib.dptr(theta), nth, _lib.dptr(phi), nph,
                      _lib.dptr(out), status.ctypes.data_as(_lib.c_ip)))
        return out, status

    def _grid_is_default(self):
        n = len(self._theta_axis)
        if self.theta.shape != (n * n,) or self.phi.shape != (n * n,):
            return False
        t, p = np.meshgrid(self._theta_axis, self._phi_axis)
        return np.array_equal(self.theta, t.ravel()) and np.array_equal(self.phi, p.ravel())

    def _trace_all(self):
        if self._grid_is_default():
            n = len(self._theta_axis)
            return self._run(self._theta_axis, n, self._phi_axis, n)
        # the caller edited .theta/.phi: trace the rays as listed
        return self._run(self.theta, len(self.theta), self.phi, 0)

    def trace_rays(self, ii):
        """Column ii of output().  The feed taper of the receiver trace is taken about np.mean(self.theta),
        np.mean(self.phi) over the WHOLE fan (rti:849-853), so a single ray cannot be traced on its own: the fan is
        traced once (one launch) and kept until .theta/.phi change."""
        c = self._cache
        if c is None or not (np.array_equal(c[0], self.theta) and np.array_equal(c[1], self.phi)):
            out, status = self._trace_all()
            c = self._cache = (np.array(self.theta, copy=True), np.array(self.phi, copy=True), out, status)
        _lib.raise_on_status(c[3][ii], strict)
        return list(c[2][:, ii])

    def output(self):
        out, status = self._trace_all()
        self.status = status
        _lib.raise_on_status(np.bitwise_or.reduce(status) if len(status) else 0, strict)
        self.out = out
        return self.out


class rx_to_lyot_model(_LyotTracer):
    """Fields on the Lyot stop ray-traced from the receiver feed (reference rti:579-885).
    output(): (20, n_pts): 0-2 Lyot hit, 3 path length, 4 feed taper, 5-7 Lyot normal,
    8-10 direction after lens 3, 11-13 hit 3B, 14-16 hit 3A, 17-19 feed position."""
    _half_width, _theta0, _phi0 = 0.28, -(np.pi / 2), (np.pi / 2)
    _rows, _entry = 20, "solat_nf_rx_trace"

    def __init__(self, p_rx, tele_geo, plotOpts):
        super().__init__(p_rx, tele_geo, plotOpts)
        self.p_rx = p_rx


class source_to_lyot_model(_LyotTracer):
    """Fields on the Lyot stop ray-traced from the holography source (reference rti:888-1382).
    output(): (26, n_pts): 0-2 Lyot hit, 3 path length, 4 amp (0 = vignetted), 5-7 normal,
    8-10 direction, 11-13 hit 2B, 14-16 hit 2A, 17-19 hit 1B, 20-22 hit 1A, 23-25 source."""
    _half_width, _theta0, _phi0 = 0.22, (np.pi / 2), (np.pi / 2)
    _rows, _entry = 26, "solat_nf_source_trace"

    def __init__(self, p_source, tele_geo, plotOpts):
        super().__init__(p_source, tele_geo, plotOpts)
        self.p_source = p_source


class aperature_fields(_LyotTracer):
    """Focal plane -> window trace through all three lenses (reference rti:33-576), (32, n_pts)."""
    _half_width, _theta0, _phi0 = 0.25, -(np.pi / 2), (np.pi / 2)
    _rows, _entry = 32, "solat_nf_toy_trace"

    def __init__(self, P_rx, tele_geo, plotOpts):
        super().__init__(P_rx, tele_geo, plotOpts)
        self.P_rx = P_rx


def _resample_index_grid(values, res_new):
    """Interpolating bicubic spline of a (res, res) field defined on linspace(-70, 70, res) in both
    axes, evaluated on linspace(-70, 70, res_new) (what interp2d(kind='cubic') did, rti:1417-1425).
    Runs on the GPU (solat_regrid_bicubic: separable not-a-knot cubic splines)."""
    res_old = values.shape[0]
    if res_old == res_new:
        return values  # an interpolating spline evaluated at its own knots
    values = _lib.as_f64(values)
    out = np.empty((res_new, res_new), dtype=np.float64)
    _lib.check(_lib.load().solat_regrid_bicubic(_lib.context().handle, _lib.dptr(values), res_old, res_new,
                                                _lib.dptr(out)))
    return out


def regrid(beam1, beam2):
    """Put both Lyot-plane fields on beam2's ray-index grid (reference rti:1385-1440).

    Like the reference it works in ray-index space on a nominal uniform +-70 mm grid and returns
    (field of beam2, field of beam1 resampled), rows 0,2 = grid X,Z, row 3 = path length, row 4 =
    amplitude.  Resampling a field onto its own grid is the identity, so only the unequal-grid case
    does any arithmetic (once per map, two small kernels)."""
    res_so = int(np.sqrt(len(beam2[0, :])))
    res_rx = int(np.sqrt(len(beam1[0, :])))
    axis = np.linspace(-70, 70, res_so)
    X2, Z2 = np.meshgrid(axis, axis)
    beam1_new = np.zeros(np.shape(beam2))
    beam2_new = np.zeros(np.shape(beam2))
    for dst, src, res in ((beam1_new, beam2, res_so), (beam2_new, beam1, res_rx)):
        dst[0, :] = np.ravel(X2)
        dst[2, :] = np.ravel(Z2)
        dst[3, :] = np.ravel(_resample_index_grid(np.reshape(src[3, :], (res, res)), res_so))
        dst[4, :] = np.ravel(_resample_index_grid(np.reshape(src[4, :], (res, res)), res_so))
    return beam1_new, beam2_new
