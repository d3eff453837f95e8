"""The product's per-ray device math (solat_optics_b200/csrc/solat_math.cuh compiled for the host by
tests/host_shadow) against the reference fixtures and against the oracle.  CPU only.  Tolerances are
the north-star ones: power 1e-6 relative, phase 1e-5 rad; per ray 1e-9 mm / 2 ulp(1e8 mm)."""
import numpy as np
import pytest

import golden_util as gu

ULP_1E8 = 1.4901161193847656e-08
P_TOL, PHI_TOL, RAY_TOL = 1e-6, 1e-5, 1e-9


def _src_axis(n):
    return np.linspace(np.pi / 2 - 0.22, np.pi / 2 + 0.22, n)


def _rx_axes(n):
    return np.linspace(-np.pi / 2 - 0.28, -np.pi / 2 + 0.28, n), np.linspace(np.pi / 2 - 0.28, np.pi / 2 + 0.28, n)


def test_nf_source_rows(shadow):
    d = gu.load("nf_so_rays")
    for i in range(int(d["n_cases"])):
        y_ap, _ = gu.nf_band(d["in_freq"][i])
        a = _src_axis(int(d["in_n_scan"][i]))
        out, st, agg = shadow.nf_source_rows([d["in_x"][i], y_ap, d["in_z"][i]], a, a)
        ref = d["out_%d" % i]
        assert agg == 0
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        assert np.array_equal(out[4] == 0, ref[4] == 0)
        np.testing.assert_allclose(np.nan_to_num(out), np.nan_to_num(ref), rtol=0, atol=RAY_TOL)


def test_nf_rx_rows(shadow):
    d = gu.load("nf_rx_rays")
    for i in range(int(d["n_cases"])):
        _, fw = gu.nf_band(d["in_freq"][i])
        th, ph = _rx_axes(int(d["in_n_scan"][i]))
        out, agg = shadow.nf_rx_rows(d["in_p_rx"][i], th, ph, fw)
        assert agg == 0
        np.testing.assert_allclose(out, d["out_%d" % i], rtol=0, atol=RAY_TOL)


def test_nf_toy_rows(shadow):
    d = gu.load("nf_toy_rays")
    for i in range(int(d["n_cases"])):
        y_ap, fw = gu.nf_band(d["in_freq"][i])
        n = int(d["in_n_scan"][i])
        th = np.linspace(-np.pi / 2 - 0.25, -np.pi / 2 + 0.25, n)
        ph = np.linspace(np.pi / 2 - 0.25, np.pi / 2 + 0.25, n)
        out, agg = shadow.nf_toy_rows(d["in_p_rx"][i], th, ph, y_ap, fw)
        assert agg == 0
        np.testing.assert_allclose(out, d["out_%d" % i], rtol=0, atol=RAY_TOL)


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("name", ["nf_pixels", "nf_map5", "nf_rx_offsets"])
def test_nf_pixels(shadow, name, packed):
    d = gu.load(name)
    for i in range(int(d["n_cases"])):
        f, n = float(d["in_freq"][i]), int(d["in_n_scan"][i])
        n_rx = int(d["in_n_scan_rx"][i]) if "in_n_scan_rx" in d else n
        if n_rx != n:
            continue
        p_rx = d["in_p_rx"][i] if "in_p_rx" in d else (0, 0, 0)
        y_ap, fw = gu.nf_band(f)
        th, ph = _rx_axes(n)
        rx, _ = shadow.nf_rx_rows(p_rx, th, ph, fw)
        a = _src_axis(n)
        z, agg = shadow.nf_pixel([d["in_x"][i], y_ap, d["in_z"][i]], a, a, rx[3], rx[4] * gu.lyot_mask(n), gu.nf_k(f),
                                packed=packed)
        ra, rp = d["amp"][i], d["phi"][i]
        assert agg == 0
        assert abs(abs(z) ** 2 - ra ** 2) <= P_TOL * ra ** 2 + 1e-300
        if ra > 0:
            assert abs(gu.wrap(np.angle(z) - rp)) <= PHI_TOL


def test_ff_rows(shadow, oracle):
    d = gu.load("ff_rays")
    for i in range(int(d["n_cases"])):
        n = int(d["in_n_scan"][i])
        th, ph = gu.ff_fan(n)
        geo = oracle.ff_geo_vec(rx_y_m=-float(d["in_offset"][i]))
        el, az = float(d["in_el"][i]), float(d["in_az"][i])
        m, a1 = shadow.ff_mirror_rows(geo, th, ph, el, az)
        f, a2 = shadow.ff_field_rows(geo, th, ph, m[6:9, 0], el, az)
        ref = d["field_%d" % i]
        assert a1 == 0 and a2 == 0
        np.testing.assert_allclose(m, d["mirror_%d" % i], rtol=0, atol=1e-9)
        assert np.array_equal(f[15] != 0, ref[15] != 0)
        np.testing.assert_allclose(f[:15], ref[:15], rtol=0, atol=1e-9)
        assert np.abs(f[15] - ref[15]).max() <= 2 * ULP_1E8


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("name", ["ff_pixels", "ff_map5", "ff_pixels_150"])
def test_ff_pixels(shadow, oracle, name, packed):
    d = gu.load(name)
    for i in range(int(d["n_cases"])):
        n = int(d["in_n_scan"][i])
        th, ph = gu.ff_fan(n)
        geo = oracle.ff_geo_vec(rx_y_m=-float(d["in_offset"][i]))
        x, y, e = gu.ff_grid(str(d["in_grid"][i]))
        z, nc, agg = shadow.ff_pixel(geo, th, ph, float(d["in_el"][i]), float(d["in_az"][i]), x, y, e,
                                     gu.ff_k(float(d["in_freq"][i])), packed=packed)
        assert agg == 0
        assert nc == int(d["n_cells"][i])
        assert abs(abs(z) ** 2 - d["amp"][i] ** 2) <= P_TOL * d["amp"][i] ** 2
        assert abs(gu.wrap(np.angle(z) - d["phi"][i])) <= PHI_TOL


def test_panel_error_matches_oracle(shadow, oracle):
    """Non-zero panel-error polynomials (saf:676-703, :820-852): several panels, some rays with no
    panel.  Only the oracle can pin this (the real pan_mod is not available: parity unpinned there)."""
    rng = np.random.default_rng(7)
    n = 14
    th, ph = gu.ff_fan(n)
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    el, az = gu.EL0 - 4 * gu.DE, 6 * gu.DE
    pan1 = np.zeros((10, 3)); pan2 = np.zeros((10, 2))
    for pm in (pan1, pan2):
        pm[0] = np.arange(pm.shape[1]) + 1; pm[1] = 1
        pm[2] = rng.normal(0, 0.02, pm.shape[1])        # a  [mm]
        pm[3:5] = rng.normal(0, 2e-5, (2, pm.shape[1]))  # b, c
        pm[5:7] = rng.normal(0, 2e-8, (2, pm.shape[1]))  # d, e
        pm[8:10] = rng.normal(0, 500, (2, pm.shape[1]))  # x0, y0
    idx1 = rng.integers(-1, 3, n * n).astype(np.int32)
    idx2 = rng.integers(-1, 2, n * n).astype(np.int32)
    m = oracle.ff_ray_mirror_pts(geo, th, ph, el, az)
    ref = oracle.ff_apert_fields(geo, th, ph, m[6:9, 0], el, az, idx1, idx2, pan1, pan2)
    out, agg = shadow.ff_field_rows(geo, th, ph, m[6:9, 0], el, az, idx1, idx2, pan1, pan2)
    assert agg == 0
    assert np.array_equal(out[15] != 0, ref[15] != 0) and (ref[15] != 0).sum() > 10
    np.testing.assert_allclose(out[:15], ref[:15], rtol=0, atol=1e-9)
    assert np.abs(out[15] - ref[15]).max() <= 3 * ULP_1E8
    # and the error terms really moved the path (otherwise this test pins nothing)
    ideal = oracle.ff_apert_fields(geo, th, ph, m[6:9, 0], el, az)
    both = (ref[15] != 0) & (ideal[15] != 0)
    assert np.abs(ref[15][both] - ideal[15][both]).max() > 1e-3


def test_panel_error_matches_reference_fixture(shadow):
    """The same per-ray source against the reference's own panel arithmetic (fixture ff_panels: synthetic 1 m panels,
    seeded errors, missing panels; only identify_panel stood in)."""
    d = gu.load("ff_panels")
    pan1, pan2 = d["pan1"], d["pan2"]
    for i in range(int(d["n_cases"])):
        th, ph = gu.ff_fan(int(d["in_n_scan"][i]))
        geo = np.array([np.arctan(0.5), -np.pi / 2 - np.arctan(1 / 3), 0, -7200.0, 0, 0, -7200.0, 1e8, -1e3 * float(d["in_offset"][i])])
        m = d["mirror_%d" % i]
        idx1, idx2 = gu.panel_indices(m, pan1, pan2)
        out, agg = shadow.ff_field_rows(geo, th, ph, m[6:9, 0], float(d["in_el"][i]), float(d["in_az"][i]), idx1, idx2, pan1, pan2)
        ref = d["field_%d" % i]
        assert agg == 0
        assert np.array_equal(out[15] != 0, ref[15] != 0)
        np.testing.assert_allclose(out[:15], ref[:15], rtol=0, atol=1e-9)
        assert np.abs(out[15] - ref[15]).max() <= 3 * ULP_1E8


def test_status_flags(shadow):
    a = _src_axis(4)
    # source behind lens 1: the 1B plane root is negative -> reference brentq ValueError -> bracket flag
    _, st, agg = shadow.nf_source_rows([0.0, -500.0, 0.0], a, a)
    assert agg & 1
    # far off-axis source: every ray vignetted, no flag raised
    out, st, agg = shadow.nf_source_rows([600.0, -1335.0, 0.0], a, a)
    assert agg == 0 and not out.any()


def test_regrid_spline_matches_fitpack(shadow):
    """The device resample (separable not-a-knot cubic splines) against FITPACK's interpolating bicubic spline,
    which is what the reference's interp2d(kind="cubic") evaluates (ray_trace_int.py:1417-1425)."""
    from scipy.interpolate import RectBivariateSpline

    for n_old, n_new in [(16, 30), (4, 9), (5, 7), (80, 150), (30, 16), (3, 5), (2, 4)]:
        ax = np.linspace(-70, 70, n_old)
        X, Z = np.meshgrid(ax, ax)
        F = 1500 + 0.01 * (X ** 2 + 0.5 * Z ** 2) + 0.3 * np.sin(X / 20) * np.cos(Z / 15)
        out = shadow.regrid(F, n_new)
        if n_old >= 4:
            new = np.linspace(-70, 70, n_new)
            ref = RectBivariateSpline(ax, ax, F.T, kx=3, ky=3, s=0)(new, new).T
            np.testing.assert_allclose(out, ref, rtol=0, atol=1e-10)
        assert np.isfinite(out).all()
        np.testing.assert_allclose(out[0, 0], F[0, 0], atol=1e-12)
        np.testing.assert_allclose(out[-1, -1], F[-1, -1], atol=1e-10)


def test_regrid_unequal_grids_known_answer(shadow):
    """Reference KAT with n_rx=16 resampled onto n_so=30 (tests/golden/nf_pixels case 12): spline resample +
    per-ray math reproduce the reference, and the answer differs from the equal-grid one at 3e-7."""
    d = gu.load("nf_pixels")
    i = [j for j in range(int(d["n_cases"])) if int(d["in_n_scan_rx"][j]) != int(d["in_n_scan"][j])][0]
    f, n, n_rx = float(d["in_freq"][i]), int(d["in_n_scan"][i]), int(d["in_n_scan_rx"][i])
    y_ap, fw = gu.nf_band(f)
    th, ph = _rx_axes(n_rx)
    rx, _ = shadow.nf_rx_rows((0, 0, 0), th, ph, fw)
    opl = shadow.regrid(rx[3].reshape(n_rx, n_rx), n).ravel()
    amp = shadow.regrid(rx[4].reshape(n_rx, n_rx), n).ravel()
    a = _src_axis(n)
    z, _ = shadow.nf_pixel([d["in_x"][i], y_ap, d["in_z"][i]], a, a, opl, amp * gu.lyot_mask(n), gu.nf_k(f))
    assert abs(abs(z) ** 2 - d["amp"][i] ** 2) <= P_TOL * d["amp"][i] ** 2
    assert abs(gu.wrap(np.angle(z) - d["phi"][i])) <= PHI_TOL


def test_ff_row_intervals_equal_the_per_ray_aperture_test(shadow, oracle):
    """The fused map's interval walk: per phi row the accepted rays must be exactly the interval the bisection finds,
    for on- and far-off-axis pixels, reversed / uneven / one-sided / dense fans, and pixels that miss the aperture."""
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    rng = np.random.default_rng(11)
    th, phi = gu.ff_fan(100)
    fans = [th, th[::-1].copy(), np.sort(th[0] + (th[-1] - th[0]) * rng.random(100)), np.linspace(np.pi + 1e-6, np.pi + 5e-5, 64),
            np.linspace(np.pi - 2e-4, np.pi + 2e-4, 257), np.linspace(np.pi - 5e-5, np.pi + 5e-5, 1000),
            np.linspace(np.pi - 3.2e-5, np.pi - 2.9e-5, 50), np.array([np.pi - 1e-5, np.pi + 2e-5])]
    pixels = [(0.0, gu.EL0), (20 * gu.DE, gu.EL0 - 20 * gu.DE), (-3e-3, 4e-3), (1e-4, -2.5e-5), (0.2, 0.1)]
    for t in fans:
        for az, el in pixels:
            bad, fallback, lo_hi = shadow.ff_row_intervals(geo, t, phi, el, az)
            assert bad == 0 and fallback == 0, (t.size, az, el, bad, fallback)
            assert (lo_hi[:, 0] <= lo_hi[:, 1]).all() and lo_hi.min() >= 0 and lo_hi.max() <= t.size
    # the default fan at the map centre: 60 % of the rays pass
    _, _, lo_hi = shadow.ff_row_intervals(geo, th, phi, gu.EL0, 0.0)
    assert int((lo_hi[:, 1] - lo_hi[:, 0]).sum()) == 6000
    # a fan that crosses the aperture plane's horizon asks for the per-ray walk
    _, fallback, _ = shadow.ff_row_intervals(geo, np.linspace(np.pi - 1.0, np.pi + 1.0, 9), phi[48:53], 1.2, 0.0)
    assert fallback > 0


def test_ff_focal_predictor_is_conservative(shadow, oracle):
    """The fused far-field map drops rays before the FP64 trace when a cubic predictor of the focal-plane position
    (fitted per pixel to 64 exactly traced rays) puts them beyond the high edge of the E_rec grid -- or below both
    low edges -- by more than 5 x its worst sample error + 2 mm.  Such a ray must be one the exact comparison
    (ffl:19) skips, resp. one that lands in cell (0, 0): checked on every accepted ray for on- and far-off-axis
    pixels, three focal offsets, small and large fans (also fans whose row length divides the sample stride), four
    grid sizes.  The worst error over ALL accepted rays must stay below HALF the margin."""
    ratio = 0.0
    for n in (100, 40, 317, 211, 1000):
        th, phi = gu.ff_fan(n)
        for offset in (1.25, 1.32, 1.4) if n < 1000 else (1.32,):
            geo = oracle.ff_geo_vec(rx_y_m=-offset)
            pixels = [(0.0, gu.EL0), (20 * gu.DE, gu.EL0 - 20 * gu.DE), (-30 * gu.DE, gu.EL0 + 30 * gu.DE), (3e-3, -2e-3), (0.02, 0.01)]
            for az, el in pixels if n < 1000 else pixels[:2]:
                for gmax in (5.0, 0.0, 20.0, 30.0) if n == 100 else (5.0,):
                    nv, worst, worst_sample, dropped, wrong = shadow.ff_pred_check(geo, th, phi, el, az, gmax, gmax)
                    assert nv > 0 and wrong == 0
                    ratio = max(ratio, worst / (5.0 * worst_sample + 2.0))
                    if gmax == 5.0 and abs(az) < 1e-3 and n >= 100:
                        assert dropped > 0.3 * nv      # the +-5 cm grid of ff_sim.py: about half of the rays miss it
                    if gmax == 30.0 and abs(az) < 1e-3:
                        assert dropped == 0
    assert ratio < 0.5, ratio          # the margin (5 x sampled error + 2 mm) covers every ray twice over
    # tiny fans: predictor off
    th, phi = gu.ff_fan(12)
    assert shadow.ff_pred_check(oracle.ff_geo_vec(rx_y_m=-1.32), th, phi, gu.EL0, 0.0, 5.0, 5.0)[0] < 0


def test_ff_ray_selection_keeps_every_winner(shadow, oracle):
    """The complete ray selection of the fused far-field map (row intervals, predictor, ff_pred_row_keep: rays beyond
    the grid dropped, of the rays that pile into cell (0, 0) one per row kept) must leave the winner tables of the pixel
    and of its mirror pixel exactly as tracing every accepted ray gives them -- for pairs and one-sided items, centred
    and shifted grids, fine and coarse grids, on- and off-axis pixels, small and large fans."""
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    saved = 0
    for n in (100, 40, 211):
        th, phi = gu.ff_fan(n)
        for az, el in [(0.0, gu.EL0), (20 * gu.DE, gu.EL0 - 20 * gu.DE), (10 * gu.DE, gu.EL0 + 15 * gu.DE), (3e-3, -2e-3)]:
            for m, g_lo, g_hi in [(51, -5.0, 5.0), (101, -5.0, 5.0), (21, -12.0, 3.0), (11, 2.0, 9.0), (61, -30.0, 30.0)]:
                for need_pos, need_neg in [(True, True), (True, False), (False, True)]:
                    if n != 100 and (m != 51 or not need_pos):
                        continue
                    on, n_acc, n_kept, diff = shadow.ff_keep_check(geo, th, phi, el, az, m, g_lo, g_hi, need_pos, need_neg)
                    assert on == 1 and diff == 0, (n, az, m, g_lo, need_pos, need_neg, n_acc, n_kept, diff)
                    assert 0 < n_kept <= n_acc
                    if (m, g_lo) == (51, -5.0) and need_pos and need_neg and n == 100:
                        saved = max(saved, n_acc - n_kept)
    assert saved > 2000          # the +-5 cm grid of ff_sim.py: a pair needs ~3100 of its 6000 accepted rays
    # large fans also go through the scout stage (FfPile: pile rays that a later, listed scout beats in every cell they
    # could reach are dropped): same winner tables, far fewer rays
    for n, cases in ((400, [(0.0, gu.EL0), (20 * gu.DE, gu.EL0 - 20 * gu.DE), (3e-3, -2e-3), (-11 * gu.DE, gu.EL0 + 7 * gu.DE)]),
                     (1000, [(5 * gu.DE, gu.EL0 - 3 * gu.DE)])):
        th, phi = gu.ff_fan(n)
        for az, el in cases:
            for m, g_lo, g_hi in [(101, -5.0, 5.0), (51, -5.0, 5.0), (21, -12.0, 3.0), (61, -30.0, 30.0)] if n == 400 else [(101, -5.0, 5.0)]:
                for need_pos, need_neg in [(True, True), (True, False), (False, True)]:
                    plain = shadow.ff_keep_check(geo, th, phi, el, az, m, g_lo, g_hi, need_pos, need_neg)
                    on, n_acc, n_kept, diff = shadow.ff_keep_check(geo, th, phi, el, az, m, g_lo, g_hi, need_pos, need_neg, pile_min=0)
                    assert on == 1 and diff == 0, (n, az, m, g_lo, need_pos, need_neg, n_acc, n_kept, diff)
                    assert 0 < n_kept <= plain[2]
                    if (m, g_lo) == (101, -5.0) and need_pos and need_neg and abs(az) < 1e-3:
                        assert n_kept < 0.75 * plain[2]


def _flip_cols(rows, nth, nph, flip_th, flip_ph):
    r = rows.reshape(rows.shape[0], nph, nth)
    if flip_th:
        r = r[:, :, ::-1]
    if flip_ph:
        r = r[:, ::-1, :]
    return r.reshape(rows.shape[0], -1)


def test_ff_ray_mirror_symmetry(shadow, oracle):
    """What the mirror-pair work items rest on: the telescope is symmetric under x -> -x, so the
    ray (theta_i, pi - phi_j) of the pixel (el, -az) is the mirror image of the ray (theta_i, phi_j) of (el, az) -
    x components change sign, everything else (the OPL above all) is equal."""
    n = 21
    th, ph = gu.ff_fan(n)
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    xs = [0, 3, 6, 9, 12]
    for el, az in ((0.0021, 0.0013), (-0.0037, 0.0044), (0.0, 0.0009)):
        mp, _ = shadow.ff_mirror_rows(geo, th, ph, el, az)
        mn, _ = shadow.ff_mirror_rows(geo, th, ph, el, -az)
        fp, a1 = shadow.ff_field_rows(geo, th, ph, mp[6:9, 0], el, az)
        fn, a2 = shadow.ff_field_rows(geo, th, ph, mn[6:9, 0], el, -az)
        assert a1 == 0 and a2 == 0
        fn = _flip_cols(fn, n, n, False, True)
        assert np.array_equal(fp[16] != 0, fn[16] != 0)
        # the reference centres each pixel on ITS ray 0 (saf:1170-1187), which is not the mirror image of the other
        # pixel's ray 0: compare absolute focal positions (the kernels carry the mirrored centre as centre_mz)
        hit = fp[16] != 0
        fp[9:12, hit] += mp[6:9, :1]
        fn[9:12, hit] += mn[6:9, :1]
        sign = np.ones(17); sign[xs] = -1
        np.testing.assert_allclose(fp[:15], sign[:15, None] * fn[:15], rtol=0, atol=1e-9)
        assert np.abs(fp[15] - fn[15]).max() <= 2 * ULP_1E8


def test_nf_ray_mirror_symmetry(shadow):
    """The optics tube is a body of revolution about y: a source at (-x, y, z) sees the ray (theta_i, pi - phi_j) as the
    mirror image of (theta_i, phi_j) from (x, y, z); likewise z with theta -> pi - theta."""
    n = 23
    a = _src_axis(n)
    y_ap, _ = gu.nf_band(150)
    for x, z in ((37.0, 0.0), (12.5, -61.0), (180.0, 95.0)):
        base, st0, _ = shadow.nf_source_rows([x, y_ap, z], a, a)
        for sx, sz in ((-1, 1), (1, -1), (-1, -1)):
            out, st, _ = shadow.nf_source_rows([sx * x, y_ap, sz * z], a, a)
            out = _flip_cols(out, n, n, sz < 0, sx < 0)
            st = _flip_cols(st[None, :], n, n, sz < 0, sx < 0)[0]
            assert np.array_equal(st, st0)
            assert np.array_equal(out[4] == 0, base[4] == 0)
            sign = np.ones(26)
            sign[[0, 8, 11, 14, 17, 20, 23]] = sx
            sign[[2, 10, 13, 16, 19, 22, 25]] = sz
            np.testing.assert_allclose(np.nan_to_num(out), sign[:, None] * np.nan_to_num(base), rtol=0, atol=RAY_TOL)


def test_lean_searchsorted_equals_numpy(shadow):
    """The branch-free searchsorted of the lean far-field binning (interpolated guess, one +-1 correction against
    the staged axis with sentinels) equals np.searchsorted(a, v, 'left') on every axis the host check admits --
    exact linspace axes, axes perturbed by up to a tenth of a step, values on, beside and far beyond the elements --
    and the host check turns down axes it must not be used on."""
    rng = np.random.default_rng(5)
    for n in (2, 3, 51, 101, 168, 1000):
        for jitter in (0.0, 1e-9, 0.1):
            a = np.linspace(-5.0, 5.0, n) if n != 51 else np.linspace(-50.0, 50.0, n)
            step = a[1] - a[0]
            if jitter:
                a = a + rng.uniform(-jitter, jitter, n) * step
                a[0], a[-1] = min(a[0], a[1] - 1e-3 * step), max(a[-1], a[-2] + 1e-3 * step)
            v = np.r_[a, np.nextafter(a, np.inf), np.nextafter(a, -np.inf), rng.uniform(a[0] - 3 * step, a[-1] + 3 * step, 4000),
                      a[0] - 1e6, a[-1] + 1e6, 0.5 * (a[:-1] + a[1:])]
            idx, lean = shadow.searchsorted_lean(a, v)
            assert lean, (n, jitter)
            np.testing.assert_array_equal(idx, np.searchsorted(a, v, "left"))
    # not sorted, not uniform enough, degenerate: the generic binning keeps those
    for bad in (np.r_[0.0, 2.0, 1.0, 3.0], np.r_[0.0, 1.0, 1.0, 3.0], np.linspace(0, 1, 50) ** 2, np.r_[1.0], np.zeros(5)):
        assert not shadow.searchsorted_lean(bad, np.r_[0.5])[1]


def test_interpolated_cell_class_is_conservative(shadow):
    """The classifier of ff_map_interp_kernel works on predicted positions u = (v - a[0]) / step.  Its contract: a
    class >= 0 is the cell np.searchsorted gives for EVERY true position within tol of the prediction; -1 means every
    such position lies beyond the last element of an axis (far_field_latrt.py:19 drops the ray); anything that a
    perturbation of up to tol could move to another cell, or in or out of the grid, is -2 (traced exactly)."""
    rng = np.random.default_rng(11)
    M, tol = 101, 3.5e-3
    ax = np.arange(M, dtype=np.float64)                      # the axis in cell units: element i sits at u = i
    n = 200000
    u = np.r_[rng.uniform(-3, M + 2, n), np.repeat(np.arange(-1, M + 1), 40) + rng.uniform(-2 * tol, 2 * tol, 40 * (M + 2)),
              M - 1 + tol * np.linspace(0.9, 1.1, 41), -tol * np.linspace(0.9, 1.1, 41)]
    v = np.r_[rng.uniform(-3, M + 2, len(u) - 1000), np.repeat(np.arange(0, 100), 10) + rng.uniform(-2 * tol, 2 * tol, 1000)]
    rng.shuffle(v)
    cls = shadow.cell_class(u, v, M, tol, tol)
    u32, v32 = u.astype(np.float32).astype(np.float64), v.astype(np.float32).astype(np.float64)

    def cells(du, dv):
        x, y = u32 + du, v32 + dv
        ix, iy = np.searchsorted(ax, x, "left"), np.searchsorted(ax, y, "left")
        return np.where((x > ax[-1]) | (y > ax[-1]), -1, iy * M + ix)           # ffl:19, ffl:22-23

    d = 0.999 * (tol - 2e-5)                                 # (single-precision rounding of u + tol itself)
    corners = [cells(su * d, sv * d) for su in (-1, 0, 1) for sv in (-1, 0, 1)]
    same = np.all([c == corners[4] for c in corners], axis=0)
    sure = cls != -2
    # whatever is classified is right under every perturbation
    for c in corners:
        np.testing.assert_array_equal(c[sure], cls[sure])
    # and the classifier is not lazy: well inside a cell (2 tol from every edge) it commits
    deep = np.all([cells(su * 2.2 * tol, sv * 2.2 * tol) == corners[4] for su in (-1, 1) for sv in (-1, 1)], axis=0)
    assert sure[deep].all() and same[sure].all()
    assert (cls[sure] == -1).sum() > 1000 and (cls[sure] >= 0).sum() > 100000 and (~sure).sum() > 1000
    # (a NaN prediction -- a lattice node that could not be traced -- never reaches the classifier: the kernel tests
    #  ux + uy == ux + uy first and traces such a ray exactly)
