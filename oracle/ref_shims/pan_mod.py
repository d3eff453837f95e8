"""Documented stand-in for the un-vendored `pan_mod` module of
McMahonCosmologyGroup/holosim-ml (unpinned; located in the reference only by the
absolute sys.path hack at scripts/ff_sim.py:6-7; call sites
solat_optics/solat_apert_field.py:10-11,21,609-617 and scripts/ff_sim.py:69-70).

Stand-in semantics (the ones SURVEY.md §8c used for its probes): ONE panel per
mirror with row=col=1, zero panel error, and every ray lies on it.  The far-field
parity of this repository is therefore pinned against the reference's own code
*plus this stand-in*; the real panel mask / panel errors are "parity unpinned".

Test infrastructure only: never imported by the product package.
"""
import numpy as np


def get_single_vert_adj_positions():
    # indexable [0]=col, [1]=row, [2]/[4]=x, [3]=y   (solat_apert_field.py:587-594)
    one = np.array([1.0])
    zero = np.array([0.0])
    adj_m1 = [one, one, zero, zero, zero]
    adj_m2 = [one, one, zero, zero, zero]
    return adj_m1, adj_m2


def panel_model_from_adjuster_offsets(mirror, err, num, save):
    # rows: [row, col, a, b, c, d, e, f, x0, y0]   (solat_apert_field.py:626,667-674)
    pm = np.zeros((10, 1))
    pm[0, 0] = 1.0
    pm[1, 0] = 1.0
    return pm


def identify_panel(x2d, y2d, x_adj, y_adj, col, row):
    # -> (2, n, n): [0]=row ids, [1]=col ids   (solat_apert_field.py:609-617)
    out = np.ones((2,) + np.shape(x2d))
    return out
