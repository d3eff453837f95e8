"""Host-side logic of the drop-in package: C-ABI surface, geometry definitions, regrid, phase unwrap,
pixel grids, sharding and the output format.  CPU only (no compute calls into the CUDA library)."""
import ctypes
import os
import re

import numpy as np
import pytest

import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "solat_b200.h")).read()
    return sorted(set(re.findall(r"SOLAT_API[^;(]*?\b(solat_[a-z0-9_]+)\s*\(", txt)))


def test_cabi_library_exports_every_declared_symbol():
    from solat_optics_b200 import _build, _lib

    _build.build()
    lib = ctypes.CDLL(_build.LIB)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libsolat_b200.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms          # the ctypes table binds exactly the header
    assert _lib.load().solat_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the product must fail loudly, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from solat_optics_b200 import _lib, maps

    with pytest.raises(_lib.SolatError):
        maps.receiver_field(maps.holo_tele_geo(140.0, 8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "solat_optics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f == "_never_", f
    for f in ("holo_sim.py", "ff_sim.py"):
        assert "oracle" not in open(os.path.join(ROOT, "scripts", f)).read()


def test_latgeo_defaults_and_setters():
    from solat_optics_b200 import ot_geo

    g = ot_geo.LatGeo()
    assert (g.n_si, g.n_vac, g.n_scan, g.lyot_y) == (3.4, 1.0, 100, 453.62)
    assert g.k == 2 * np.pi / ((30.0 / 150.0) * 0.01)
    g.set_wavelength(0.002); g.set_nscan(40); g.set_fwhp(0.3); g.set_res(1e-4)
    assert (g.k, g.n_scan, g.th_fwhp, g.de_ang) == (2 * np.pi / 0.002, 40, 0.3, 1e-4)
    assert ot_geo.LENS3_Y == 180.8 - ((2.5 + 1.513) - (35.16 - 32.19))


def test_asphere_slopes_are_gradients_of_sags():
    from solat_optics_b200 import ot_geo

    h = 1e-4
    for z, dz in ((ot_geo.z1a, ot_geo.d_z1a), (ot_geo.z2a, ot_geo.d_z2a), (ot_geo.z3a, ot_geo.d_z3a), (ot_geo.z3b, ot_geo.d_z3b)):
        for x, y in ((30.0, -50.0), (0.5, 120.0), (-80.0, 15.0)):
            gx, gy = dz(x, y)
            assert abs(gx - (z(x + h, y) - z(x - h, y)) / (2 * h)) < 1e-8
            assert abs(gy - (z(x, y + h) - z(x, y - h)) / (2 * h)) < 1e-8


def test_ot_frames_roundtrip_and_layout():
    from solat_optics_b200 import ot_geo as g

    p = (12.0, -300.0, 45.0)
    for fwd, back in ((g.tele_into_m1b, g.m1b_into_tele), (g.tele_into_m1a, g.m1a_into_tele), (g.tele_into_m2b, g.m2b_into_tele),
                      (g.tele_into_m2a, g.m2a_into_tele), (g.tele_into_m3a, g.m3a_into_tele), (g.tele_into_m3b, g.m3b_into_tele),
                      (g.tele_into_lyot, g.lyot_into_tele)):
        np.testing.assert_allclose(back(*fwd(*p)), p, atol=1e-10)
    # th=-pi/2 -> local = (x, -z, y+Y0); th=+pi/2 (1B) -> (x, z, -(y+Y0))
    np.testing.assert_allclose(g.tele_into_m2b(*p), (12.0, -45.0, -300.0 + 592.76), atol=1e-10)
    np.testing.assert_allclose(g.tele_into_m1b(*p), (12.0, 45.0, -(-300.0 + 975.08)), atol=1e-10)


def test_lat_mirror_polynomials_and_frames():
    from solat_optics_b200 import latrt_geo as lg

    x, y = 812.0, -1430.0
    z = sum(lg.a1[i, j] * (x / 3000) ** i * (y / 3000) ** j for i in range(7) for j in range(7))
    assert abs(lg.z1(x, y) - z) < 1e-12
    h = 1e-3
    for zf, dzf in ((lg.z1, lg.d_z1), (lg.z2, lg.d_z2)):
        gx, gy = dzf(x, y)
        assert abs(gx - (zf(x + h, y) - zf(x - h, y)) / (2 * h)) < 1e-9
        assert abs(gy - (zf(x, y + h) - zf(x, y - h)) / (2 * h)) < 1e-9
    el, az = 3e-4, -7e-4
    p = (100.0, -7000.0, 4000.0)
    q = lg.rotate_az_el(*p, el, az)
    # rotation about the centre keeps the distance to it
    c = np.array([0, -7200.0, 0])
    assert abs(np.linalg.norm(np.array(q) - c) - np.linalg.norm(np.array(p) - c)) < 1e-9
    np.testing.assert_allclose(lg.tele_into_m1(*lg.m1_into_tele(5.0, 6.0, 7.0), 0.0, 0.0), (5.0, 6.0, 7.0), atol=1e-9)
    np.testing.assert_allclose(lg.tele_into_m2(*lg.m2_into_tele(5.0, 6.0, 7.0), 0.0, 0.0), (5.0, 6.0, 7.0), atol=1e-9)
    g = lg.initialize_telescope_geometry()
    assert g.th2 == (-np.pi / 2) - np.arctan(1 / 3) and g.y_rotc == -7.2 and g.N_scan == 100


def test_snell_vec_matches_kernel_form():
    from solat_optics_b200 import ray_trace_int as rti

    n = np.array([0.1, 0.9, -0.2]); n /= np.linalg.norm(n)
    s = np.array([0.05, -0.98, 0.1]); s /= np.linalg.norm(s)
    mu = 1 / 3.4
    c = -np.dot(n, s)
    expect = mu * s + (mu * c - np.sqrt(1 - mu * mu * (1 - c * c))) * n
    np.testing.assert_allclose(rti.snell_vec(1.0, 3.4, n, s), expect, atol=1e-15)
    assert np.isnan(rti.snell_vec(3.4, 1.0, np.array([0, 0, 1.0]), np.array([0.6, 0, -0.8]))).all()   # TIR


def test_regrid_identity_and_layout():
    from solat_optics_b200 import ray_trace_int as rti

    rng = np.random.default_rng(0)
    rx = rng.normal(size=(20, 144)); so = rng.normal(size=(26, 144))
    b1, b2 = rti.regrid(rx, so)
    assert b1.shape == so.shape and b2.shape == so.shape
    ax = np.linspace(-70, 70, 12)
    X, Z = np.meshgrid(ax, ax)
    np.testing.assert_array_equal(b1[0], X.ravel()); np.testing.assert_array_equal(b2[2], Z.ravel())
    np.testing.assert_array_equal(b1[3], so[3]); np.testing.assert_array_equal(b1[4], so[4])   # source field
    np.testing.assert_array_equal(b2[3], rx[3]); np.testing.assert_array_equal(b2[4], rx[4])   # rx field


def test_do_unwrap_matches_reference_bit_for_bit():
    from solat_optics_b200 import phase_correct

    d = gu.load("unwrap")
    out = phase_correct.do_unwrap(d["phi_deg"].copy())
    np.testing.assert_array_equal(out, d["unwrapped_deg"])


def test_pixel_grids_and_band_switch():
    from solat_optics_b200 import maps

    px = maps.holo_pixels(51, 100, 0)
    rx = np.linspace(-50, 50, 51)
    assert px.shape == (2601, 2)
    assert tuple(px[51 * 3 + 7]) == (rx[3], rx[7])          # row index i*len+j, column 0 slowest (holo_sim.py:95-108)
    lo, hi = maps.holo_tele_geo(90.0), maps.holo_tele_geo(140.0)
    assert (lo.y_ap, hi.y_ap) == (-1278.0, -1335.0)
    assert lo.th_fwhp == 40.5 * np.pi / 180 and hi.th_fwhp == 25 * np.pi / 180
    assert hi.k == 2 * np.pi / (0.299792458 / 140.0)
    tg = maps.ff_tele_geo(140, 1.32)
    assert tg.k == 2 * np.pi / ((30.0 / 140) * 0.01) and tg.rx_y == -1.32 and tg.el0 == np.arctan(7.2 / 100e3)
    fp = maps.ff_pixels(tg, 0.25, 20)
    assert fp.shape == (1681, 2)
    np.testing.assert_allclose(fp[41 * 5 + 9], (-20 * gu.DE + 5 * gu.DE, gu.EL0 - 20 * gu.DE + 9 * gu.DE), atol=1e-18)


def test_shard_bounds_cover_everything_once():
    from solat_optics_b200 import maps

    for n in (0, 1, 7, 1681, 2601, 160801):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = maps.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi)) if n < 5000 else [(lo, hi)]
            if n < 5000:
                assert seen == list(range(n))
            else:
                assert seen[0][0] == 0 and seen[-1][1] == n and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))


def test_output_txt_format(tmp_path):
    """4 columns, %.18e, space separated, no header, row index i*len+j (holo_sim.py:131-139); the shipped
    reference output output_files/ff_1.32m_140GHz_highres.txt has the same shape for 41x41."""
    from solat_optics_b200 import maps

    tg = maps.ff_tele_geo(140, 1.32)
    px = maps.ff_pixels(tg, 0.25, 20)
    amp = np.arange(len(px), dtype=float); pha = -amp
    p = tmp_path / "out" / "ff.txt"
    maps.save_map_txt(str(p), px[:, 0], px[:, 1], amp, pha)
    lines = open(p).read().splitlines()
    assert len(lines) == 1681 and all(len(l.split(" ")) == 4 for l in lines)
    assert re.fullmatch(r"-?\d\.\d{18}e[+-]\d{2}", lines[0].split(" ")[0])
    back = np.loadtxt(p)
    np.testing.assert_array_equal(back[:, 2], amp)
    # az spans +-20*0.25' about 0, el about 7.2e-5 -- the coordinates of the reference's shipped file
    assert abs(back[0, 0] + 1.4544410433e-3) < 1e-12 and abs(back[0, 1] - (7.2e-5 - 1.4544410433e-3)) < 1e-11


def test_far_field_grid_matches_shipped_reference_output():
    """Coordinates of the one result file the reference ships (output_files/ff_1.32m_140GHz_highres.txt, produced by
    `ff_sim.py 140 0.25 20 <offset>`): rows 0, 1, 41, 840, 1680 of its (az, el) columns, copied here as known
    answers.  Its amplitude/phase columns cannot be reproduced (near-field input and pan_mod are not shipped)."""
    from solat_optics_b200 import maps

    shipped = {0: (-0.001454441043328608, -0.001382441043453024), 1: (-0.001454441043328608, -0.0013097189912865937),
               41: (-0.0013817189911621776, -0.001382441043453024), 840: (0.0, 7.199999987558391e-05),
               1680: (0.001454441043328608, 0.0015264410432041919)}
    px = maps.ff_pixels(maps.ff_tele_geo(140, 1.32), 0.25, 20)
    assert len(px) == 1681
    for row, (az, el) in shipped.items():
        assert px[row, 0] == az and abs(px[row, 1] - el) <= 1e-18


def test_near_field_table_to_e_rec():
    from solat_optics_b200 import maps

    px = maps.holo_pixels(5, 100, 0)
    amp = 1 + px[:, 0] ** 2 / 1e4
    pha = gu.wrap(0.002 * (px[:, 0] ** 2 + px[:, 1] ** 2))
    x, y, e = maps.near_field_to_e_rec(np.c_[px[:, 0], px[:, 1], amp, pha])
    assert x.shape == (5, 5)
    np.testing.assert_allclose(x[0, :], np.linspace(-5, 5, 5))     # ff_sim.py:55: x <- column 1, cm
    np.testing.assert_allclose(y[:, 0], np.linspace(-5, 5, 5))
    np.testing.assert_allclose(np.abs(e), amp.reshape(5, 5))


def test_mirror_pair_order_keeps_partners_adjacent():
    """Multi-rank far-field maps shard a pixel list ordered so that (+az, el) sits next to (-az, el); the order is a
    permutation, tolerant of the 1e-19 rad asymmetry of np.linspace, and leaves az = 0 pixels alone."""
    from solat_optics_b200 import maps

    tg = maps.ff_tele_geo(140, 1.32)
    px = maps.ff_pixels(tg, 0.25, 20)
    order = maps.mirror_pair_order(px)
    assert sorted(order.tolist()) == list(range(len(px)))
    q = px[order]
    n0 = int((q[:, 0] == 0).sum())
    assert n0 == 41 and (q[:n0, 0] == 0).all()
    rest = q[n0:]
    assert len(rest) % 2 == 0
    a, b = rest[0::2], rest[1::2]
    assert (a[:, 0] > 0).all() and (b[:, 0] < 0).all()
    np.testing.assert_array_equal(a[:, 1], b[:, 1])
    np.testing.assert_allclose(a[:, 0], -b[:, 0], rtol=0, atol=1e-17)
    # near field: the four positions (+-x, +-z) of a group are neighbours
    px = maps.holo_pixels(res=11, sweep=100)
    order = maps.mirror_pair_order(px, fold_second=True)
    assert sorted(order.tolist()) == list(range(len(px)))
    q = np.abs(px[order])
    key = np.round(q * 1e6)
    change = 1 + np.nonzero((np.diff(key, axis=0) != 0).any(axis=1))[0]
    groups = np.split(np.arange(len(px)), change)
    assert len(groups) == 6 * 6 and max(len(g) for g in groups) == 4


def test_panel_layout_host_semantics():
    """solat_apert_field.PanelLayout: half-open rectangles, later panels override, ids (0, 0) outside, first-match
    model columns (saf:626, :805-809)."""
    from solat_optics_b200 import solat_apert_field as saf

    lay = saf.PanelLayout([0, 5, -10], [10, 15, -5], [0, 5, -10], [10, 15, -5], [1, 2, 3], [4, 5, 6])
    xs = np.array([[0.0, 9.999, 10.0], [7.0, 12.0, -7.0]]); ys = np.array([[0.0, 9.999, 10.0], [7.0, 12.0, -7.0]])
    ids = lay.identify_panel(xs, ys)
    np.testing.assert_array_equal(ids[0], [[1, 2, 2], [2, 2, 3]])
    np.testing.assert_array_equal(ids[1], [[4, 5, 5], [5, 5, 6]])
    assert (lay.identify_panel(np.array([20.0]), np.array([20.0])) == 0).all()
    pm = np.zeros((10, 4)); pm[0] = [9, 2, 2, 1]; pm[1] = [9, 5, 5, 4]
    np.testing.assert_array_equal(lay.model_columns(pm), [3, 1, -1])
    adj = saf.PanelLayout.from_adjusters([0.0, 100.0], [0.0, 0.0], [1, 2], [7, 7], 90.0, 50.0)
    np.testing.assert_allclose(adj.rect, [[-45, 45, -25, 25], [55, 145, -25, 25]])
    one, two = saf.panel_backend().panel_layouts()
    assert (one.identify_panel(np.array([[123.0]]), np.array([[-4e3]])) == 1).all()
