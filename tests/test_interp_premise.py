"""The numerical premise of ff_map_interp_kernel, checked on the CPU with the oracle: over a 1000^2
fan, the focal-plane position of a ray is reproduced by cubic (4-point Lagrange) interpolation, in both fan directions,
between exactly traced rays 16 apart to well within the probe threshold of the kernel (1e-3 mm; its ambiguity margin
is 3e-3 mm, the E_rec cells of BASELINE config 5 are 1 mm)."""
import numpy as np
import pytest


def _lagrange_weights(f, s):
    t = 1.0 + f / float(s)
    return np.stack([-(t - 1) * (t - 2) * (t - 3) / 6, t * (t - 2) * (t - 3) / 2, -t * (t - 1) * (t - 3) / 2, t * (t - 1) * (t - 2) / 6], -1)


@pytest.mark.parametrize("corner", [(0, 0), (400, 0), (200, 200), (37, 390)])
def test_lattice_interpolation_predicts_focal_positions(oracle, corner):
    import bench
    from solat_optics_b200 import maps

    n, s = 1000, 16
    tg = maps.ff_tele_geo(bench.FREQ, bench.OFFSET)
    az, el = maps.ff_pixels(tg, bench.RES_ARCMIN, bench.NUM).reshape(401, 401, 2)[corner]
    th, ph = maps.ff_fan(n)
    out = oracle.ff_apert_fields(oracle.ff_geo_vec(rx_y_m=-bench.OFFSET), th, ph, np.zeros(3), el, az)
    ok = (out[15] != 0).reshape(n, n)                      # [j (phi)][i (theta)]
    worst = 0.0
    for row in (9, 11):                                    # focal-plane x and z [mm]
        f = out[row].reshape(n, n)
        nodes, node_ok = f[::s, ::s], ok[::s, ::s]         # lattice [m][k] = ray (i = 16 k, j = 16 m)
        jj, ii = np.nonzero(ok)
        mb, kb = jj // s, ii // s
        inside = (mb >= 1) & (kb >= 1) & (mb + 2 < nodes.shape[0]) & (kb + 2 < nodes.shape[1])
        jj, ii, mb, kb = jj[inside], ii[inside], mb[inside], kb[inside]
        wj, wi = _lagrange_weights(jj - mb * s, s), _lagrange_weights(ii - kb * s, s)
        pred = np.zeros(len(jj))
        good = np.ones(len(jj), dtype=bool)
        for a in range(4):
            for b in range(4):
                pred += wj[:, a] * wi[:, b] * nodes[mb - 1 + a, kb - 1 + b]
                good &= node_ok[mb - 1 + a, kb - 1 + b]
        assert good.sum() > 400000
        worst = max(worst, float(np.abs(pred[good] - f[jj[good], ii[good]]).max()))
    assert worst < 5e-4, worst                             # measured 2.3e-4 .. 3.1e-4 mm; the kernel's probes allow 1e-3 mm


def test_winners_from_interpolated_cells_equal_exact_winners(oracle, shadow):
    """The selection logic of ff_map_interp_kernel on the CPU, with the kernel's own classifier (host shadow): every
    accepted ray of a 1000^2 fan gets its cell from the single-precision lattice prediction when ff_cell_class commits,
    from its exact position when it does not (the kernel traces such a ray), and the latest ray per cell must be the
    one the exact positions give (far_field_latrt.py:19-28) -- in the pixel's table and in a mirror-like table (x
    negated, z shifted, rows counted backwards)."""
    import bench
    from solat_optics_b200 import maps

    n, s, M = 1000, 16, 101
    tg = maps.ff_tele_geo(bench.FREQ, bench.OFFSET)
    az, el = maps.ff_pixels(tg, bench.RES_ARCMIN, bench.NUM).reshape(401, 401, 2)[310, 95]
    th, ph = maps.ff_fan(n)
    out = oracle.ff_apert_fields(oracle.ff_geo_vec(rx_y_m=-bench.OFFSET), th, ph, np.zeros(3), el, az)
    ok = (out[15] != 0).reshape(n, n)
    gx = np.linspace(-5.0, 5.0, M)                           # the E_rec axes of config 5 [cm]
    inv = (M - 1) / (gx[-1] - gx[0])
    cx, cz = np.median(out[9][ok.ravel()]), np.median(out[11][ok.ravel()])       # the grid sits mid-beam, as in the map
    x_cm, z_cm = (out[9].reshape(n, n) - cx) / 10.0, (out[11].reshape(n, n) - cz) / 10.0
    tol = np.float32(3.0e-4 * inv + 5.0e-4)
    jj, ii = np.nonzero(ok)
    mb, kb = jj // s, ii // s
    wj, wi = _lagrange_weights(jj - mb * s, s).astype(np.float32), _lagrange_weights(ii - kb * s, s).astype(np.float32)
    for mirror in (False, True):
        X, Z = (-x_cm, z_cm + 0.037) if mirror else (x_cm, z_cm)
        key = ((n - 1 - jj) if mirror else jj) * n + ii + 1
        # exact table
        xe, ze = X[jj, ii], Z[jj, ii]
        keep = ~((xe > gx[-1]) | (ze > gx[-1]))
        cell_e = np.searchsorted(gx, ze, "left") * M + np.searchsorted(gx, xe, "left")
        exact = np.zeros(M * M, dtype=np.int64)
        np.maximum.at(exact, cell_e[keep], key[keep])
        # prediction from the lattice (single precision, cell units); nodes outside the aperture or the fan: no prediction
        ux = ((X - gx[0]) * inv).astype(np.float32)
        uy = ((Z - gx[0]) * inv).astype(np.float32)
        nx, ny, nok = ux[::s, ::s], uy[::s, ::s], ok[::s, ::s]
        inside = (mb >= 1) & (kb >= 1) & (mb + 2 < nx.shape[0]) & (kb + 2 < nx.shape[1])
        px, py = np.zeros(len(jj), dtype=np.float32), np.zeros(len(jj), dtype=np.float32)
        good = inside.copy()
        m1, k1 = np.where(inside, mb, 1), np.where(inside, kb, 1)
        for a in range(4):
            rx, ry = np.zeros(len(jj), dtype=np.float32), np.zeros(len(jj), dtype=np.float32)
            for b in range(4):
                rx += wi[:, b] * nx[m1 - 1 + a, k1 - 1 + b]
                ry += wi[:, b] * ny[m1 - 1 + a, k1 - 1 + b]
                good &= nok[m1 - 1 + a, k1 - 1 + b]
            px += wj[:, a] * rx
            py += wj[:, a] * ry
        cls = shadow.cell_class(px, py, M, float(tol), float(tol))
        cls = np.where(good, cls, -2)
        assert np.abs(px[good] - ux[jj, ii][good]).max() < 1e-3 and np.abs(py[good] - uy[jj, ii][good]).max() < 1e-3   # the probes' bound
        cell = np.where(cls >= 0, cls, np.where((cls == -2) & keep, cell_e, -1))     # ambiguous: the exact trace decides
        got = np.zeros(M * M, dtype=np.int64)
        sel = cell >= 0
        np.maximum.at(got, cell[sel], key[sel])
        np.testing.assert_array_equal(got, exact)
        frac_amb = float(((cls == -2) & good).mean())
        assert (exact > 0).sum() == M * M and 0.0 < frac_amb < 0.02, frac_amb
