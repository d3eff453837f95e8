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
