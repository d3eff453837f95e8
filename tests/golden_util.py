"""Helpers shared by the tests: golden-fixture loading and the input conventions of
oracle/gen_golden.py (which produced tests/golden/*.npz from the unmodified reference)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DE = 0.25 / 60 * np.pi / 180           # scripts/ff_sim.py:94 with res = 0.25 arcmin
EL0 = float(np.arctan(7.2 / 100e3))    # scripts/ff_sim.py:77-79
C_HOLO = 0.299792458                   # scripts/holo_sim.py:56


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def nf_band(freq):
    """scripts/holo_sim.py:40-45 -> (y_ap [mm], th_fwhp [rad])"""
    if freq <= 120:
        return -1213 - (6.5 * 10), 40.5 * np.pi / 180
    return -1213 - (12.2 * 10), 25 * np.pi / 180


def nf_k(freq):
    return 2 * np.pi / (C_HOLO / freq)   # holo_sim.py:56-57 [rad/m]


def ff_k(freq):
    return 2 * np.pi / ((30.0 / freq) * 0.01)   # ff_sim.py:43-44


def ff_fan(n):
    return np.linspace(np.pi - 0.00005, np.pi + 0.00005, n), np.linspace(0, np.pi, n)   # ff_sim.py:65-66


def ff_grid(kind):
    """Same E_rec grids as oracle/gen_golden.py:ff_grid ("cfg5" = bench.e_rec_config5)."""
    if kind == "nf51":
        ax = np.linspace(-50, 50, 51) / 1e1
    elif kind == "cfg5":
        ax = np.linspace(-5, 5, 101)
    else:
        ax = np.linspace(-30, 30, 61)
    y, x = np.meshgrid(ax, ax, indexing="ij")
    sig = 1.5 if kind in ("nf51", "cfg5") else 9.0
    amp = np.exp(-(x ** 2 + y ** 2) / (2 * sig ** 2))
    ph = 0.3 * (x ** 2 + y ** 2) * (1.0 if kind in ("nf51", "cfg5") else 0.03)
    return x, y, amp * np.exp(1j * ph)


def wrap(a):
    return (np.asarray(a) + np.pi) % (2 * np.pi) - np.pi


def lyot_mask(n_scan, r_lyot=136.54 / 2):
    """scripts/holo_sim.py:168-171 on regrid's uniform index-space grid (rti:1413-1415)."""
    ax = np.linspace(-70, 70, n_scan)
    X2, Z2 = np.meshgrid(ax, ax)
    return np.sqrt(np.ravel(X2) ** 2 + np.ravel(Z2) ** 2) <= r_lyot


Y_CENT_M1 = -7201.003729431267   # solat_apert_field.py:19


def panel_identify(x2d, y2d, x_adj=None, y_adj=None, col=None, row=None):
    """Same synthetic panel layout as oracle/gen_golden.py:panel_identify (1 m x 1 m panels): (2, n, n) row, col ids."""
    return np.stack([np.floor(np.asarray(y2d) / 1000.0) + 10, np.floor(np.asarray(x2d) / 1000.0) + 10])


def panel_columns(panel_model, ids):
    """First column of the panel model whose (row, col) equals the ray's ids, -1 if none (saf:626, :805-809)."""
    rows, cols = np.ravel(ids[0]), np.ravel(ids[1])
    out = np.full(rows.size, -1, dtype=np.int32)
    for idx in range(panel_model.shape[1] - 1, -1, -1):
        out[(rows == panel_model[0, idx]) & (cols == panel_model[1, idx])] = idx
    return out


def panel_indices(mirror, pan1, pan2):
    """Per-ray panel-model columns the way saf:596-621 derives them from ray_mirror_pts' output."""
    n = int(round(np.sqrt(mirror.shape[1])))
    sh = (n, n)
    id2 = panel_identify(mirror[0].reshape(sh), mirror[2].reshape(sh))
    id1 = panel_identify(mirror[3].reshape(sh), mirror[4].reshape(sh) - Y_CENT_M1)
    return panel_columns(pan1, id1), panel_columns(pan2, id2)
