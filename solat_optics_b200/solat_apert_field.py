"""Far-field LAT ray tracers: drop-in for the reference's `solat_optics/solat_apert_field.py`.

`ray_mirror_pts`, `aperature_fields_from_panel_model` and the class twin `RayMirrorPts` keep the
reference signatures and output layouts (saf:307-567 -> (9, N^2); saf:570-1009 -> (17, N^2);
saf:24-304).  The per-ray Python/brentq loops become one CUDA launch each.

Panel model plug point.  The reference takes panel identification and panel-error polynomials
from the external module `pan_mod` (McMahonCosmologyGroup/holosim-ml, not vendored, unpinned;
saf:10-11,21,609-617).  Here:
  * if a `pan_mod` module is importable (or registered with `set_panel_backend`), its
    `get_single_vert_adj_positions` / `identify_panel` are called on the host exactly where the
    reference calls them, and the resulting per-ray panel ids are matched against the panel models;
  * otherwise a documented stand-in is used: one panel (row = col = 1) covering the whole mirror.
The panel-error polynomial itself (a + b x + c y + d (x^2+y^2) + e x y, saf:689-703,838-852) is
evaluated inside the kernel for whatever coefficients the panel models carry.
"""
import ctypes

import numpy as np

from . import _lib
from . import latrt_geo  # noqa: F401
from .latrt_geo import *  # noqa: F401,F403
from .latrt_geo import initialize_telescope_geometry

th2 = (-np.pi / 2) - initialize_telescope_geometry.th_2
y_cent_m1 = -7201.003729431267  # centre of the LAT coordinates [mm] (reference saf:19)

strict = False
_backend = None
_backend_probed = False


class _SinglePanel:
    """Stand-in panel backend: one panel with ids (1, 1) over the whole mirror."""

    @staticmethod
    def get_single_vert_adj_positions():
        one, zero = np.array([1.0]), np.array([0.0])
        return [one, one, zero, zero, zero], [one, one, zero, zero, zero]

    @staticmethod
    def identify_panel(x2d, y2d, x_adj, y_adj, col, row):
        return np.ones((2,) + np.shape(x2d))

    @staticmethod
    def panel_model_from_adjuster_offsets(mirror, err, num, save):
        pm = np.zeros((10, 1))
        pm[0, 0] = pm[1, 0] = 1.0
        return pm

    @staticmethod
    def panel_layouts():
        big = 1e9
        one = PanelLayout([-big], [big], [-big], [big], [1.0], [1.0])
        return one, one


class PanelLayout:
    """Rectangular panel layout of one mirror, the form in which panel identification runs on the GPU.

    Panel k covers [x0[k], x1[k]) x [y0[k], y1[k]) in the coordinates the reference hands to `identify_panel`
    (saf:596-612: M1 -> (x, y - y_cent_m1), M2 -> (x, z) of the ideal-mirror hit, telescope frame, mm) and carries the ids
    (row[k], col[k]); a later panel overrides an earlier one; a point in no panel gets ids (0, 0) -- what an
    `identify_panel` that loops over the panels and writes their ids into zero-initialised arrays returns.
    A panel backend that can describe itself this way exposes `panel_layouts() -> (layout_m1, layout_m2)`; maps with
    panel models then never leave the GPU.  The layout of holosim-ml's own pan_mod is not part of the reference:
    "parity unpinned" (DESIGN.md); `from_adjusters` builds the obvious candidate (one rectangle per adjuster-defined
    panel centre)."""

    def __init__(self, x0, x1, y0, y1, row, col):
        self.rect = np.ascontiguousarray(np.stack([np.asarray(v, dtype=np.float64).ravel() for v in (x0, x1, y0, y1)], axis=1))
        self.row = np.asarray(row, dtype=np.float64).ravel()
        self.col = np.asarray(col, dtype=np.float64).ravel()
        if not (len(self.rect) == len(self.row) == len(self.col)):
            raise ValueError("x0, x1, y0, y1, row, col must have one entry per panel")

    @classmethod
    def from_adjusters(cls, x_cent, y_cent, col, row, size_x, size_y):
        """One size_x x size_y rectangle centred on every (x_cent, y_cent) (e.g. the mean adjuster position of a panel)."""
        x_cent, y_cent = np.asarray(x_cent, dtype=np.float64), np.asarray(y_cent, dtype=np.float64)
        return cls(x_cent - size_x / 2, x_cent + size_x / 2, y_cent - size_y / 2, y_cent + size_y / 2, row, col)

    def identify_panel(self, x2d, y2d, *_unused):
        """Host evaluation with the call shape of pan_mod.identify_panel: (2, ...) row ids, col ids."""
        x2d, y2d = np.asarray(x2d, dtype=np.float64), np.asarray(y2d, dtype=np.float64)
        out = np.zeros((2,) + x2d.shape)
        for k in range(len(self.rect)):
            x0, x1, y0, y1 = self.rect[k]
            m = (x2d >= x0) & (x2d < x1) & (y2d >= y0) & (y2d < y1)
            out[0][m] = self.row[k]
            out[1][m] = self.col[k]
        return out

    def model_columns(self, panel_model):
        """Per panel: first column of `panel_model` with the panel's (row, col) ids, -1 if none (saf:626, :805-809)."""
        return _first_match_index(panel_model, self.row, self.col)


def set_panel_backend(module):
    """Register the object providing identify_panel / get_single_vert_adj_positions (e.g. holosim-ml's
    pan_mod).  None restores the single-panel stand-in."""
    global _backend, _backend_probed
    _backend = module
    _backend_probed = True


def panel_backend():
    global _backend, _backend_probed
    if not _backend_probed:
        _backend_probed = True
        try:
            import pan_mod  # noqa: F401  (external holosim-ml, if the user has it on sys.path)

            _backend = pan_mod
        except ImportError:
            _backend = None
    return _backend if _backend is not None else _SinglePanel


def _call_rows(entry, tele_geo, theta, phi, el, az, extra):
    lib = _lib.load()
    ctx = _lib.context()
    geom = _lib.ff_geom(tele_geo)
    theta = _lib.as_f64(theta)
    phi = _lib.as_f64(phi)
    n = theta.size * phi.size
    rows = 9 if entry == "solat_ff_ray_mirror_pts" else 17
    out = np.empty((rows, n), dtype=np.float64)
    status = np.empty(n, dtype=np.int32)
    fn = getattr(lib, entry)
    args = [ctx.handle, ctypes.byref(geom), _lib.dptr(theta), theta.size, _lib.dptr(phi), phi.size] + extra + [
        float(el), float(az), _lib.dptr(out), status.ctypes.data_as(_lib.c_ip)]
    _lib.check(fn(*args))
    _lib.raise_on_status(np.bitwise_or.reduce(status) if n else 0, strict)
    return out


def ray_mirror_pts(tele_geo, theta, phi, el, az):
    """Source -> aperture plane -> M1 -> M2 hit points for the theta x phi fan (reference saf:307-567).
    Returns (9, N^2): rows 0-2 M2 hit, 3-5 M1 hit, [6:9, 0] focal-plane centre."""
    return _call_rows("solat_ff_ray_mirror_pts", tele_geo, theta, phi, el, az, [])


def _first_match_index(panel_model, rows, cols):
    """Vectorised `np.where((pm[0]==row) & (pm[1]==col))[0][0]` per ray, -1 where nothing matches."""
    panel_model = np.asarray(panel_model, dtype=np.float64)
    first = {}
    for idx in range(panel_model.shape[1] - 1, -1, -1):
        first[(panel_model[0, idx], panel_model[1, idx])] = idx
    return np.fromiter((first.get((r, c), -1) for r, c in zip(rows, cols)), dtype=np.int32, count=len(rows))


def aperature_fields_from_panel_model(panel_model1, panel_model2, tele_geo, theta, phi, rxmirror, el, az):
    """Source -> aperture -> M1 (+panel error) -> M2 (+panel error) -> focal plane with path length
    (reference saf:570-1009).  Returns (17, N^2): 0-2 M2 hit, 3-5 M1 hit, 6-8 M2 normal, 9-11 focal
    hit minus focal centre, 12-14 outgoing direction, 15 path length, 16 amplitude."""
    backend = panel_backend()
    adj_pos_m1, adj_pos_m2 = backend.get_single_vert_adj_positions()
    n_phi = len(np.atleast_1d(phi))
    n_side = int(np.sqrt(len(np.atleast_1d(theta)) * n_phi))
    rxmirror = np.asarray(rxmirror, dtype=np.float64)
    shape = (n_side, n_side)
    # host side of saf:587-621: which panel does every ray sit on
    pan_id_m2 = backend.identify_panel(np.reshape(rxmirror[0, :], shape), np.reshape(rxmirror[2, :], shape),
                                       adj_pos_m2[4], adj_pos_m2[3], adj_pos_m2[0], adj_pos_m2[1])
    pan_id_m1 = backend.identify_panel(np.reshape(rxmirror[3, :], shape), np.reshape(rxmirror[4, :], shape) - y_cent_m1,
                                       adj_pos_m1[2], adj_pos_m1[3], adj_pos_m1[0], adj_pos_m1[1])
    idx1 = _first_match_index(panel_model1, np.ravel(pan_id_m1[0]), np.ravel(pan_id_m1[1]))
    idx2 = _first_match_index(panel_model2, np.ravel(pan_id_m2[0]), np.ravel(pan_id_m2[1]))
    pm1 = _lib.as_f64(panel_model1)
    pm2 = _lib.as_f64(panel_model2)
    centre = _lib.as_f64(rxmirror[6:9, 0])
    extra = [_lib.dptr(centre), idx1.ctypes.data_as(_lib.c_ip), idx2.ctypes.data_as(_lib.c_ip),
             _lib.dptr(pm1), pm1.shape[1], _lib.dptr(pm2), pm2.shape[1]]
    return _call_rows("solat_ff_apert_fields", tele_geo, theta, phi, el, az, extra)


class RayMirrorPts:
    """Class form of ray_mirror_pts (reference saf:24-304): `.output()` returns the (9, N^2) array.
    The reference forks an mp.Pool for n_pts > 3000; here it is one kernel launch either way."""

    def __init__(self, tele_geo, theta, phi, el, az):
        self.tele_geo = tele_geo
        self._theta_axis = np.asarray(theta, dtype=np.float64)
        self._phi_axis = np.asarray(phi, dtype=np.float64)
        t, p = np.meshgrid(self._theta_axis, self._phi_axis)
        self.theta = np.ravel(t)
        self.phi = np.ravel(p)
        self.el = el
        self.az = az
        self.n_pts = len(self.theta)
        self.out = np.zeros((9, self.n_pts))

    def output(self):
        self.out = ray_mirror_pts(self.tele_geo, self._theta_axis, self._phi_axis, self.el, self.az)
        return self.out
