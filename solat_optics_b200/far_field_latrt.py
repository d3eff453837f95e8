"""Drop-in for the reference's `solat_optics/far_field_latrt.py`."""
import numpy as np

from . import _lib


def project_to_data(sim, data_x, data_y, tele_geo):
    """Scatter exp(i k L/1e3) of every valid far-field ray into the near-field data grid
    (reference far_field_latrt.py:9-32): nearest upper bin by searchsorted, rays beyond the grid
    maximum skipped, later rays overwrite earlier ones.  One CUDA scatter (atomicMax on the ray
    index) + fill instead of the reference's Python loop.  Returns complex128 (M, M)."""
    sim = _lib.as_f64(sim)
    data_x = _lib.as_f64(np.real(data_x))
    data_y = _lib.as_f64(np.real(data_y))
    if sim.ndim != 2 or sim.shape[0] < 16:
        raise ValueError("sim must be the (17, n_pts) array of aperature_fields_from_panel_model")
    if sim.shape[0] != 17:
        sim = _lib.as_f64(np.vstack([sim, np.zeros((17 - sim.shape[0], sim.shape[1]))])[:17])
    m = data_x.shape[0]
    if data_x.shape != (m, m) or data_y.shape != (m, m):
        raise ValueError("data_x and data_y must be square grids of equal shape")
    out = np.zeros((m, m), dtype=np.complex128)
    ctx = _lib.context()
    _lib.check(_lib.load().solat_ff_project(ctx.handle, _lib.dptr(sim), sim.shape[1], _lib.dptr(data_x),
                                            _lib.dptr(data_y), m, float(tele_geo.k),
                                            out.ctypes.data_as(_lib.c_dp)))
    return out
