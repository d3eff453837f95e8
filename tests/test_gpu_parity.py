"""Parity tests proper: the CUDA path, called through the reference-shaped Python API (which binds the
C ABI of include/solat_b200.h), against the golden fixtures of the unmodified reference and against the
CPU oracle.  Tolerances (BASELINE.json north_star): power 1e-6 relative, phase 1e-5 rad per pixel;
diagnostics per ray: 1e-9 mm, far-field path length 2 ulp of 1e8 mm."""
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu

ULP_1E8 = 1.4901161193847656e-08
P_TOL, PHI_TOL, RAY_TOL = 1e-6, 1e-5, 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class PlotOpts:
    plot = 0


def _assert_pixels(amp, phase, ref_amp, ref_phase, max_bad=0):
    amp, phase, ref_amp, ref_phase = map(np.ravel, (amp, phase, ref_amp, ref_phase))
    bad_p = np.abs(amp ** 2 - ref_amp ** 2) > P_TOL * ref_amp ** 2 + 1e-300
    bad_f = (np.abs(gu.wrap(phase - ref_phase)) > PHI_TOL) & (ref_amp > 0)
    bad = bad_p | bad_f
    assert bad.sum() <= max_bad, "pixels out of tolerance: %s" % np.flatnonzero(bad)[:10]
    return int(bad.sum())


# ------------------------------------------------------------------------------------------------ near field
def _nf_geo(freq, n):
    from solat_optics_b200 import maps

    return maps.holo_tele_geo(float(freq), int(n))


def test_library_is_the_cuda_one():
    from solat_optics_b200 import _lib

    ctx = _lib.context(0)
    assert ctx.cc[0] == 10 and ctx.sm_count >= 100
    before = _lib.launch_count()
    assert ctx.fp64_peak_tflops(2) > 1.0
    assert _lib.launch_count() > before


def test_fast_math_primitives():
    """Device rcp / rsqrt / sincos building blocks against libm (the accuracy the Newton schedule relies on)."""
    from solat_optics_b200 import _lib

    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(0.5, 2.0, 5000), 10 ** rng.uniform(-4, 9, 5000)])
    o = _lib.selftest_math(xs)
    assert np.max(np.abs(o[1] * xs - 1)) < 5e-12          # one Newton step: Newton step sizes
    assert np.max(np.abs(o[2] * xs - 1)) < 4e-16          # two steps: geometry
    assert np.max(np.abs(o[4] * np.sqrt(xs) - 1)) < 6e-16
    big = rng.uniform(2.8e8, 3.0e8, 5000)                 # far-field phases k L / 1e3
    o = _lib.selftest_math(big)
    assert np.max(np.abs(o[6] - np.cos(big))) < 2e-15 and np.max(np.abs(o[7] - np.sin(big))) < 2e-15
    mid = rng.uniform(-3e4, 3e4, 20000)                   # near-field phases
    o = _lib.selftest_math(mid)
    assert np.max(np.abs(o[8] - np.cos(mid))) < 1e-10 and np.max(np.abs(o[9] - np.sin(mid))) < 1e-10


def test_nf_rx_rows():
    from solat_optics_b200 import ray_trace_int as rti

    d = gu.load("nf_rx_rays")
    for i in range(int(d["n_cases"])):
        tg = _nf_geo(d["in_freq"][i], d["in_n_scan"][i])
        m = rti.rx_to_lyot_model(list(d["in_p_rx"][i]), tg, PlotOpts)
        out = m.output()
        assert out.shape == (20, m.n_pts) and out.flags["C_CONTIGUOUS"] and out is m.out
        np.testing.assert_allclose(out, d["out_%d" % i], rtol=0, atol=RAY_TOL)
        if i == 0:
            # the feed taper (row 4) is taken about the means of the whole fan (rti:849-853): one ray == its column
            for ii in (0, 3, m.n_pts - 1):
                np.testing.assert_array_equal(np.array(m.trace_rays(ii)), out[:, ii])
            assert len({m.trace_rays(ii)[4] for ii in (0, m.n_pts // 2 + 1)}) == 2


def test_nf_source_rows_and_trace_rays():
    from solat_optics_b200 import ray_trace_int as rti

    d = gu.load("nf_so_rays")
    for i in range(int(d["n_cases"])):
        tg = _nf_geo(d["in_freq"][i], d["in_n_scan"][i])
        m = rti.source_to_lyot_model([d["in_x"][i], tg.y_ap, d["in_z"][i]], tg, PlotOpts)
        out = m.output()
        ref = d["out_%d" % i]
        assert out.shape == (26, m.n_pts)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        assert np.array_equal(out[4] == 0, ref[4] == 0)
        np.testing.assert_allclose(np.nan_to_num(out), np.nan_to_num(ref), rtol=0, atol=RAY_TOL)
        if i == 1:
            for ii in (0, 7, m.n_pts - 1):
                one = np.array(m.trace_rays(ii))
                np.testing.assert_array_equal(np.nan_to_num(one), np.nan_to_num(out[:, ii]))


def test_nf_toy_rows():
    from solat_optics_b200 import ray_trace_int as rti

    d = gu.load("nf_toy_rays")
    for i in range(int(d["n_cases"])):
        tg = _nf_geo(d["in_freq"][i], d["in_n_scan"][i])
        out = rti.aperature_fields(list(d["in_p_rx"][i]), tg, PlotOpts).output()
        np.testing.assert_allclose(out, d["out_%d" % i], rtol=0, atol=RAY_TOL)


def test_nf_edited_ray_list_uses_paired_mode():
    from solat_optics_b200 import ray_trace_int as rti

    tg = _nf_geo(140.0, 6)
    m = rti.source_to_lyot_model([3.0, tg.y_ap, -8.0], tg, PlotOpts)
    full = m.output().copy()
    keep = np.array([5, 1, 30, 17])
    m.theta, m.phi = m.theta[keep], m.phi[keep]
    sub = m.output()
    np.testing.assert_array_equal(np.nan_to_num(sub), np.nan_to_num(full[:, keep]))


def test_nf_strict_mode_raises_like_brentq():
    from solat_optics_b200 import ray_trace_int as rti

    tg = _nf_geo(140.0, 4)
    m = rti.source_to_lyot_model([0.0, -500.0, 0.0], tg, PlotOpts)    # behind lens 1: no root in [5, 14000]
    rti.strict = True
    try:
        with pytest.raises(ValueError):
            m.output()
    finally:
        rti.strict = False
    m.output()
    assert (m.status & 1).any()


def _nf_pixel_gpu(freq, n, n_rx, p_rx, x, z):
    from solat_optics_b200 import maps

    tg = _nf_geo(freq, n)
    rx_opl, rx_w = maps.receiver_field(tg, tuple(p_rx), n_rx, n)
    plan = maps.NearFieldPlan(tg, [tg.k], rx_opl, rx_w, n, device=0)
    amp, ph, st = plan.run([[x, z]])
    plan.close()
    return amp[0, 0], ph[0, 0], st


@pytest.mark.parametrize("name", ["nf_pixels", "nf_rx_offsets"])
def test_nf_pixel_known_answers(name):
    d = gu.load(name)
    amps, phs = [], []
    for i in range(int(d["n_cases"])):
        n_rx = int(d["in_n_scan_rx"][i]) if "in_n_scan_rx" in d else int(d["in_n_scan"][i])
        a, p, st = _nf_pixel_gpu(d["in_freq"][i], int(d["in_n_scan"][i]), n_rx, d["in_p_rx"][i], d["in_x"][i], d["in_z"][i])
        assert st == 0
        amps.append(a); phs.append(p)
    _assert_pixels(np.array(amps), np.array(phs), d["amp"], d["phi"])
    fully_vignetted = d["amp"] == 0
    assert np.all(np.array(amps)[fully_vignetted] == 0)
    if name == "nf_pixels":
        assert fully_vignetted.any()


def test_nf_small_map_and_frequency_sweep(oracle):
    from solat_optics_b200 import maps

    d = gu.load("nf_map5")
    px, amp, ph = maps.holo_map(140.0, res=5, sweep=100, n_scan=20, device=0)
    np.testing.assert_array_equal(px, np.c_[d["in_x"], d["in_z"]])
    _assert_pixels(amp[0], ph[0], d["amp"], d["phi"])
    # three frequencies of one band share one trace (BASELINE config 3) == three separate maps, bit for bit
    px, amp3, ph3 = maps.holo_map([150.0, 220.0, 280.0], res=5, n_scan=20, device=0)
    for i, f in enumerate((150.0, 220.0, 280.0)):
        _, a1, p1 = maps.holo_map(f, res=5, n_scan=20, device=0)
        np.testing.assert_array_equal(amp3[i], a1[0]); np.testing.assert_array_equal(ph3[i], p1[0])
    with pytest.raises(ValueError):
        maps.holo_map([90.0, 150.0], res=3, n_scan=8, device=0)          # different bands: y_ap differs


def test_nf_receiver_offset_sweep_shares_the_source_trace(oracle):
    """BASELINE config 4 (quant_systematics): several receiver positions in one plan == one plan per position,
    bit for bit, with and without several frequencies; checked against the oracle for an offset feed."""
    from solat_optics_b200 import maps

    offsets = [(x, 0.0, z) for x in (-60.0, 0.0, 60.0) for z in (-60.0, 0.0, 60.0)] + [(30.0, 0.0, -15.0)]   # 10 tables
    px, amp, ph = maps.receiver_offset_sweep(150.0, offsets, res=5, n_scan=24, device=0)
    assert amp.shape == (10, 25)
    tg = maps.holo_tele_geo(150.0, 24)
    for j in (0, 4, 9):
        rx_opl, rx_w = maps.receiver_field(tg, offsets[j])
        a1, p1, st = maps.NearFieldPlan(tg, [tg.k], rx_opl, rx_w, device=0).run(px)
        assert st == 0
        np.testing.assert_array_equal(a1[0], amp[j]); np.testing.assert_array_equal(p1[0], ph[j])
        orx = oracle.nf_rx_trace(offsets[j], tg.th_fwhp, 24)
        oa, op = oracle.nf_map(px, tg.y_ap, 24, orx[3], orx[4] * gu.lyot_mask(24), [tg.k])
        _assert_pixels(amp[j], ph[j], oa, op)
    # 3 tables x 2 wavenumbers in one plan
    tabs = [maps.receiver_field(tg, o) for o in offsets[:3]]
    k2 = [tg.k, maps.holo_tele_geo(220.0, 24).k]
    plan = maps.NearFieldPlan(tg, k2, np.stack([t[0] for t in tabs]), np.stack([t[1] for t in tabs]), device=0)
    a, p, st = plan.run(px)
    assert a.shape == (3, 2, 25) and plan.channels == 6 and st == 0
    for j in range(3):
        a1, p1, _ = maps.NearFieldPlan(tg, k2, tabs[j][0], tabs[j][1], device=0).run(px)
        np.testing.assert_array_equal(a1, a[j]); np.testing.assert_array_equal(p1, p[j])


def test_nf_many_table_sweep_contraction(oracle, monkeypatch):
    """More than 16 receiver tables on a plan with both mirror symmetries: the quad items are traced once into
    exp(i k L) and contracted against every table (nf_phase_kernel + nf_contract_kernel).  Against one plan per table
    (the fused kernel; another summation order: 1e-11 of the map's peak), against the fused kernel on the same plan
    (SOLAT_NF_CONTRACT=0), against the oracle; bit-identical under permutation, subsets and lists above 4096 pixels."""
    from solat_optics_b200 import maps

    tg = maps.holo_tele_geo(150.0, 36)
    ax = np.linspace(-60.0, 60.0, 5)
    offsets = [(x, 0.0, z) for x in ax for z in ax][:21]
    opl, w = maps.receiver_fields(tg, offsets, device=0)
    px = maps.holo_pixels(res=9, sweep=100)
    k2 = [tg.k, 1.31 * tg.k]
    for kk in ([tg.k], k2):
        plan = maps.NearFieldPlan(tg, kk, opl, w, device=0)
        a, p, st = plan.run(px)                                   # [21, nfreq, 81]
        assert st == 0 and a.shape == (21, len(kk), 81)
        monkeypatch.setenv("SOLAT_NF_CONTRACT", "0")
        a_f, p_f, _ = plan.run(px)
        monkeypatch.delenv("SOLAT_NF_CONTRACT")
        e, e_f = a * np.exp(1j * p), a_f * np.exp(1j * p_f)
        assert np.abs(e - e_f).max() <= 1e-11 * np.abs(e_f).max()
        for j in (0, 7, 20):
            a1, p1, _ = maps.NearFieldPlan(tg, kk, opl[j], w[j], device=0).run(px)
            e1 = a1 * np.exp(1j * p1)
            assert np.abs(e[j] - e1).max() <= 1e-11 * np.abs(e1).max()
            orx = oracle.nf_rx_trace(offsets[j], tg.th_fwhp, 36)
            oa, op = oracle.nf_map(px, tg.y_ap, 36, orx[3], orx[4] * gu.lyot_mask(36), kk)
            for f in range(len(kk)):
                _assert_pixels(a[j, f], p[j, f], oa[f], op[f])
        rng = np.random.default_rng(5)
        for sel in (rng.permutation(len(px)), np.nonzero(px[:, 1] < 0)[0], rng.permutation(len(px))[:29],
                    np.tile(np.arange(len(px)), 52)[:4200]):
            a_s, p_s, _ = plan.run(px[sel])
            np.testing.assert_array_equal(a_s, a[:, :, sel]); np.testing.assert_array_equal(p_s, p[:, :, sel])
        plan.close()


def test_nf_mirror_pairs(oracle, monkeypatch):
    """The optics tube is symmetric under x -> -x and so is the launch fan: the near-field map traces a ray once for
    the source positions (+x, z) and (-x, z).  Canonical pixels (x >= 0, z >= 0) bit-identical with the pair walk on and
    off, the others (always computed as mirror images) within 1e-9 and within tolerance of the oracle; shuffled, one-sided and
    duplicated lists give the same values."""
    from solat_optics_b200 import maps

    tg = maps.holo_tele_geo(150.0, 36)
    rx_opl, rx_w = maps.receiver_field(tg, (20.0, 0.0, -35.0))          # an off-axis receiver: asymmetric weights
    px = maps.holo_pixels(res=9, sweep=100)
    kk = [tg.k, maps.holo_tele_geo(220.0, 36).k, maps.holo_tele_geo(280.0, 36).k]
    plan = maps.NearFieldPlan(tg, kk, rx_opl, rx_w, device=0)
    a_on, p_on, st = plan.run(px)
    assert st == 0
    monkeypatch.setenv("SOLAT_NF_PAIRS", "0")
    a_off, p_off, st = maps.NearFieldPlan(tg, kk, rx_opl, rx_w, device=0).run(px)
    monkeypatch.delenv("SOLAT_NF_PAIRS")
    pos = (px[:, 0] >= 0) & (px[:, 1] >= 0)              # the fan is symmetric in theta too: z < 0 pixels are mirror images as well
    np.testing.assert_array_equal(a_on[:, pos], a_off[:, pos])
    np.testing.assert_array_equal(p_on[:, pos], p_off[:, pos])
    np.testing.assert_allclose(a_on, a_off, rtol=1e-9, atol=0)
    assert np.abs(gu.wrap(p_on - p_off)).max() < 1e-9
    rx = oracle.nf_rx_trace((20.0, 0.0, -35.0), tg.th_fwhp, 36)
    oa, op = oracle.nf_map(px, tg.y_ap, 36, rx[3], rx[4] * gu.lyot_mask(36), kk)
    for f in range(3):
        _assert_pixels(a_on[f], p_on[f], oa[f], op[f])
    rng = np.random.default_rng(4)
    perm = rng.permutation(len(px))
    a_p, p_p, _ = plan.run(px[perm])
    np.testing.assert_array_equal(a_p, a_on[:, perm]); np.testing.assert_array_equal(p_p, p_on[:, perm])
    a_n, p_n, _ = plan.run(px[~pos])
    np.testing.assert_allclose(a_n, a_on[:, ~pos], rtol=1e-12, atol=0)
    dup = np.concatenate([px[:11], px[:11]])
    a_d, _, _ = plan.run(dup)
    np.testing.assert_allclose(a_d, np.concatenate([a_on[:, :11], a_on[:, :11]], axis=1), rtol=1e-12, atol=0)


@pytest.mark.parametrize("nfreq", [1, 2, 4])
def test_nf_mirror_quads(oracle, monkeypatch, nfreq):
    """With a fan symmetric in theta as well the work item is the source (|x|, |z|) and feeds the four pixels (+-x, +-z).  Every pixel
    is bit-identical to the pair walk (SOLAT_NF_QUADS=0) and to the one-pixel-per-CTA walk with mirrors, whatever subset
    or order of the list is asked for; against no mirroring at all (SOLAT_NF_PAIRS=0) and the oracle within tolerance."""
    from solat_optics_b200 import maps

    tg = maps.holo_tele_geo(150.0, 36)
    rx_opl, rx_w = maps.receiver_field(tg, (20.0, 0.0, -35.0))
    px = maps.holo_pixels(res=9, sweep=100)
    kk = [tg.k, maps.holo_tele_geo(220.0, 36).k, 1.31 * tg.k, 0.77 * tg.k][:nfreq]
    plan = maps.NearFieldPlan(tg, kk, rx_opl, rx_w, device=0)
    a_q, p_q, st = plan.run(px)
    assert st == 0
    monkeypatch.setenv("SOLAT_NF_QUADS", "0")
    a_2, p_2, _ = maps.NearFieldPlan(tg, kk, rx_opl, rx_w, device=0).run(px)
    monkeypatch.delenv("SOLAT_NF_QUADS")
    np.testing.assert_array_equal(a_q, a_2); np.testing.assert_array_equal(p_q, p_2)
    monkeypatch.setenv("SOLAT_NF_PAIRS", "0")
    a_0, p_0, _ = maps.NearFieldPlan(tg, kk, rx_opl, rx_w, device=0).run(px)
    monkeypatch.delenv("SOLAT_NF_PAIRS")
    canon = (px[:, 0] >= 0) & (px[:, 1] >= 0)
    np.testing.assert_array_equal(a_q[:, canon], a_0[:, canon])
    np.testing.assert_allclose(a_q, a_0, rtol=1e-9, atol=0)
    assert np.abs(gu.wrap(p_q - p_0)).max() < 1e-9
    rx = oracle.nf_rx_trace((20.0, 0.0, -35.0), tg.th_fwhp, 36)
    oa, op = oracle.nf_map(px, tg.y_ap, 36, rx[3], rx[4] * gu.lyot_mask(36), kk)
    for f in range(nfreq):
        _assert_pixels(a_q[f], p_q[f], oa[f], op[f])
    rng = np.random.default_rng(11)
    subsets = [rng.permutation(len(px)), np.nonzero(px[:, 1] < 0)[0], np.nonzero((px[:, 0] < 0) & (px[:, 1] > 0))[0],
               np.nonzero(px[:, 0] * px[:, 1] > 0)[0], rng.permutation(len(px))[:37], np.concatenate([np.arange(9), np.arange(9)])]
    # a list above 4096 positions is not paired (one pixel per CTA, mirror images all the same)
    subsets.append(np.tile(np.arange(len(px)), 52)[:4200])
    for sel in subsets:
        a_s, p_s, _ = plan.run(px[sel])
        np.testing.assert_array_equal(a_s, a_q[:, sel]); np.testing.assert_array_equal(p_s, p_q[:, sel])


def test_nf_batched_receiver_fields_equal_the_per_receiver_ones():
    """solat_nf_rx_fields (all receiver positions of a sweep in one launch, only the path-length and taper rows)
    against rx_to_lyot_model(...).output() + regrid + mask, receiver by receiver: bit for bit."""
    from solat_optics_b200 import maps

    tg = maps.holo_tele_geo(150.0, 37)
    pos = [(0.0, 0.0, 0.0), (60.0, 0.0, -30.0), (-25.0, 3.0, 100.0)]
    opl, w = maps.receiver_fields(tg, pos, device=0)
    for k, p in enumerate(pos):
        o1, w1 = maps.receiver_field(tg, p)
        np.testing.assert_array_equal(opl[k], o1)
        np.testing.assert_array_equal(w[k], w1)


def test_nf_full_default_map_against_oracle(oracle):
    """BASELINE config 1 at full size: 51x51 pixels x 100^2 rays, every pixel against the oracle."""
    from solat_optics_b200 import maps

    px, amp, ph = maps.holo_map(140.0, device=0)
    assert amp.shape == (1, 2601)
    tg = maps.holo_tele_geo(140.0, 100)
    rx = oracle.nf_rx_trace([0, 0, 0], tg.th_fwhp, 100)
    oa, op = oracle.nf_map(px, tg.y_ap, 100, rx[3], rx[4] * gu.lyot_mask(100), [tg.k])
    _assert_pixels(amp, ph, oa, op)
    # the two n_scan=100 reference KATs sit on this grid: (0,0) is pixel 25*51+25
    d = gu.load("nf_pixels")
    _assert_pixels(amp[0, 25 * 51 + 25], ph[0, 25 * 51 + 25], d["amp"][0], d["phi"][0])


def test_nf_properties_at_full_size():
    """Size-independent properties on the full config-1 workload: (a) splitting the pixel list across calls
    (what N ranks do) is bit-identical to one call; (b) the device-pointer entry equals the host entry;
    (c) linear in the receiver weight; (d) mirror symmetry x -> -x of the optics tube."""
    import torch

    from solat_optics_b200 import maps

    tg = maps.holo_tele_geo(140.0, 100)
    rx_opl, rx_w = maps.receiver_field(tg)
    plan = maps.NearFieldPlan(tg, [tg.k], rx_opl, rx_w, device=0)
    px = maps.holo_pixels()
    amp, ph, st = plan.run(px)
    assert st == 0 and plan.active_rays == int(np.count_nonzero(rx_w))
    a1, p1, _ = plan.run(px[:1301]); a2, p2, _ = plan.run(px[1301:])
    np.testing.assert_array_equal(np.c_[a1, a2], amp); np.testing.assert_array_equal(np.c_[p1, p2], ph)
    d_src = torch.from_numpy(px).cuda(); d_amp = torch.empty(2601, dtype=torch.float64, device="cuda"); d_ph = torch.empty_like(d_amp)
    plan.run_device(d_src, d_amp, d_ph, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_amp.cpu().numpy(), amp[0]); np.testing.assert_array_equal(d_ph.cpu().numpy(), ph[0])
    assert plan.status() == 0
    plan2 = maps.NearFieldPlan(tg, [tg.k], rx_opl, 2.0 * rx_w, device=0)
    b, q, _ = plan2.run(px[::50])
    np.testing.assert_allclose(b, 2 * amp[:, ::50], rtol=1e-13); np.testing.assert_allclose(q, ph[:, ::50], atol=1e-12)
    grid_a = amp[0].reshape(51, 51); grid_p = ph[0].reshape(51, 51)
    np.testing.assert_allclose(grid_a[::-1], grid_a, rtol=1e-9)       # x -> -x
    assert np.abs(gu.wrap(grid_p[::-1] - grid_p)).max() < 1e-9


# ------------------------------------------------------------------------------------------------ far field
def _ff_geo(freq, offset):
    from solat_optics_b200 import maps

    return maps.ff_tele_geo(int(freq), float(offset))


def test_ff_rows():
    from solat_optics_b200 import solat_apert_field as saf

    d = gu.load("ff_rays")
    pm = saf.panel_backend().panel_model_from_adjuster_offsets(1, np.zeros(1092), 1, 0)
    for i in range(int(d["n_cases"])):
        tg = _ff_geo(d["in_freq"][i], d["in_offset"][i])
        th, ph = gu.ff_fan(int(d["in_n_scan"][i]))
        el, az = float(d["in_el"][i]), float(d["in_az"][i])
        m = saf.ray_mirror_pts(tg, th, ph, el, az)
        np.testing.assert_allclose(m, d["mirror_%d" % i], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(saf.RayMirrorPts(tg, th, ph, el, az).output(), m)
        f = saf.aperature_fields_from_panel_model(pm, pm, tg, th, ph, m, el, az)
        ref = d["field_%d" % i]
        assert f.shape == ref.shape and np.array_equal(f[15] != 0, ref[15] != 0)
        np.testing.assert_allclose(f[:15], ref[:15], rtol=0, atol=1e-9)
        assert np.abs(f[15] - ref[15]).max() <= 2 * ULP_1E8
        np.testing.assert_array_equal(f[16], ref[16])


def test_ff_panel_errors_against_oracle(oracle):
    from solat_optics_b200 import solat_apert_field as saf

    rng = np.random.default_rng(7)
    n = 14
    th, ph = gu.ff_fan(n)
    tg = _ff_geo(140, 1.32)
    el, az = gu.EL0 - 4 * gu.DE, 6 * gu.DE
    pan1 = np.zeros((10, 3)); pan2 = np.zeros((10, 2))
    for pm in (pan1, pan2):
        pm[0] = np.arange(pm.shape[1]) + 1; pm[1] = 1
        pm[2] = rng.normal(0, 0.02, pm.shape[1]); pm[3:5] = rng.normal(0, 2e-5, (2, pm.shape[1]))
        pm[5:7] = rng.normal(0, 2e-8, (2, pm.shape[1])); pm[8:10] = rng.normal(0, 500, (2, pm.shape[1]))
    rows1 = rng.integers(0, 4, (n, n)).astype(float); rows2 = rng.integers(0, 3, (n, n)).astype(float)   # 0 = no panel

    class Backend:
        calls = 0

        @staticmethod
        def get_single_vert_adj_positions():
            z = np.zeros(1)
            return [z, z, z, z, z], [z, z, z, z, z]

        @classmethod
        def identify_panel(cls, x2d, y2d, x_adj, y_adj, col, row):
            cls.calls += 1
            r = rows2 if cls.calls == 1 else rows1            # reference calls M2 first (saf:609-612)
            return np.stack([r, np.ones_like(r)])

    saf.set_panel_backend(Backend)
    try:
        m = saf.ray_mirror_pts(tg, th, ph, el, az)
        out = saf.aperature_fields_from_panel_model(pan1, pan2, tg, th, ph, m, el, az)
    finally:
        saf.set_panel_backend(None)
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    idx1 = rows1.ravel().astype(np.int32) - 1; idx2 = rows2.ravel().astype(np.int32) - 1
    ref = oracle.ff_apert_fields(geo, th, ph, m[6:9, 0], el, az, idx1, idx2, pan1, pan2)
    assert np.array_equal(out[15] != 0, ref[15] != 0) and (ref[15] != 0).sum() > 10
    np.testing.assert_allclose(out[:15], ref[:15], rtol=0, atol=1e-9)
    assert np.abs(out[15] - ref[15]).max() <= 3 * ULP_1E8


def test_ff_panel_model_map(oracle):
    """Fused far-field map with panel models (ff_sim.py:204-221 with non-zero adjuster errors): mirror coordinates on
    the GPU -> identify_panel on the host -> fused map.  Pinned on the reference fixture ff_panels (synthetic 1 m
    panels, only identify_panel stood in), then a larger fan against the oracle's per-pixel composition, then the
    error-free single-panel layout against the plain fused map."""
    from solat_optics_b200 import maps
    from solat_optics_b200 import solat_apert_field as saf

    class Backend:
        @staticmethod
        def get_single_vert_adj_positions():
            z = np.zeros(1)
            return [z, z, z, z, z], [z, z, z, z, z]

        identify_panel = staticmethod(gu.panel_identify)

    d = gu.load("ff_panels")
    pan1, pan2 = d["pan1"], d["pan2"]
    n = int(d["in_n_scan"][0])
    th, ph = gu.ff_fan(n)
    x, y, e = gu.ff_grid("wide61")
    tgf = maps.ff_tele_geo(140, 1.32)
    azel = np.c_[d["in_az"], d["in_el"]]
    plan = maps.FarFieldPlan(tgf, th, ph, x, y, e, device=0)
    xy, st = plan.mirror_xy(azel)
    assert st == 0
    for i in range(len(azel)):
        m = d["mirror_%d" % i]
        np.testing.assert_allclose(xy[i], m[[0, 2, 3, 4]], rtol=0, atol=1e-9)
    a, p, st = plan.run_panel_models(azel, pan1, pan2, backend=Backend, batch=2)
    assert st == 0
    _assert_pixels(a, p, d["amp"], d["phi"])
    # larger fan, more pixels: the oracle's ray_mirror_pts -> apert_fields -> project -> sum, pixel by pixel
    n = 40
    th, ph = gu.ff_fan(n)
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    azel = np.array([[0.3 * gu.DE, gu.EL0], [11 * gu.DE, gu.EL0 - 7 * gu.DE], [-20 * gu.DE, gu.EL0 + 20 * gu.DE]])
    plan = maps.FarFieldPlan(tgf, th, ph, x, y, e, device=0)
    a, p, st = plan.run_panel_models(azel, pan1, pan2, backend=Backend)
    assert st == 0
    plain, _, _ = plan.run(azel)
    for i, (az, el) in enumerate(azel):
        m = oracle.ff_ray_mirror_pts(geo, th, ph, el, az)
        idx1, idx2 = gu.panel_indices(m, pan1, pan2)
        f = oracle.ff_apert_fields(geo, th, ph, m[6:9, 0], el, az, idx1, idx2, pan1, pan2)
        beam = np.sum(oracle.ff_project(f, x, y, tgf.k) * e) / len(x) ** 2
        _assert_pixels(a[i:i + 1], p[i:i + 1], [abs(beam)], [np.angle(beam)])
        assert abs(a[i] - plain[i]) > 1e-4 * plain[i]          # the panel errors and gaps really changed the pixel
    # error-free single panel = the plain fused map
    saf.set_panel_backend(None)
    pm = saf.panel_backend().panel_model_from_adjuster_offsets(1, np.zeros(1092), 1, 0)
    a0, p0, st = plan.run_panel_models(azel, pm, pm)
    _assert_pixels(a0, p0, plain, plan.run(azel)[1])
    # through maps.ff_map (the CLI path)
    px, am, phs = maps.ff_map(140, 0.25, 1, 1.32, x, y, e, n_scan=n, device=0, panel_models=(pm, pm))
    px2, am2, phs2 = maps.ff_map(140, 0.25, 1, 1.32, x, y, e, n_scan=n, device=0)
    _assert_pixels(am, phs, am2, phs2)


def test_ff_project_to_data(oracle):
    from solat_optics_b200 import far_field_latrt as ffl
    from solat_optics_b200 import solat_apert_field as saf

    tg = _ff_geo(140, 1.32)
    pm = saf.panel_backend().panel_model_from_adjuster_offsets(1, None, 1, 0)
    for kind in ("nf51", "wide61"):
        x, y, e = gu.ff_grid(kind)
        th, ph = gu.ff_fan(40)
        el, az = gu.EL0 + 2 * gu.DE, -5 * gu.DE
        m = saf.ray_mirror_pts(tg, th, ph, el, az)
        f = saf.aperature_fields_from_panel_model(pm, pm, tg, th, ph, m, el, az)
        out = ffl.project_to_data(f, x, y, tg)
        ref = oracle.ff_project(f, x, y, tg.k)                   # same sim array in: cells must agree exactly
        assert out.dtype == np.complex128 and out.shape == x.shape
        assert np.array_equal(out != 0, ref != 0) and np.count_nonzero(ref) > 50
        np.testing.assert_allclose(out, ref, rtol=0, atol=1e-9)   # same k*L/1e3 argument; reduction error only


@pytest.mark.parametrize("name", ["ff_pixels", "ff_map5", "ff_pixels_150", "ff_pixels_big"])
def test_ff_pixel_known_answers(name):
    from solat_optics_b200 import maps

    d = gu.load(name)
    amps, phs = [], []
    for i in range(int(d["n_cases"])):
        tg = _ff_geo(d["in_freq"][i], d["in_offset"][i])
        th, ph = gu.ff_fan(int(d["in_n_scan"][i]))
        x, y, e = gu.ff_grid(str(d["in_grid"][i]))
        plan = maps.FarFieldPlan(tg, th, ph, x, y, e, device=0)
        a, p, st = plan.run([[d["in_az"][i], d["in_el"][i]]])
        plan.close()
        assert st == 0
        amps.append(a[0]); phs.append(p[0])
    _assert_pixels(np.array(amps), np.array(phs), d["amp"], d["phi"])


def test_ff_full_map_against_oracle(oracle):
    """BASELINE config 2 geometry at full size: ff_sim.py 140 0.25 20 1.32 -> 41x41 pixels x 100^2 rays, every
    pixel against the oracle.  A ray within ~1e-8 mm of a bin edge may land in the neighbouring cell
    (SURVEY.md section 7 'last-write-wins'); at most 2 such pixels are tolerated and none is expected."""
    from solat_optics_b200 import maps

    x, y, e = gu.ff_grid("wide61")
    px, amp, ph = maps.ff_map(140, 0.25, 20, 1.32, x, y, e, device=0)
    assert amp.shape == (1681,)
    tg = maps.ff_tele_geo(140, 1.32)
    th, phi = maps.ff_fan(100)
    oa, op = oracle.ff_map(oracle.ff_geo_vec(rx_y_m=-1.32), px, th, phi, x, y, e, tg.k)
    _assert_pixels(amp, ph, oa, op, max_bad=2)


def test_ff_properties_at_full_size():
    """(a) sharded == unsharded bit for bit; (b) device entry == host entry; (c) linear in E_rec;
    (d) a global phase on E_rec rotates the result; (e) more than one batch per pixel (n_scan=100) and
    a ragged last batch give the same winners as a single-batch fan restricted to the same rays."""
    import torch

    from solat_optics_b200 import maps

    x, y, e = gu.ff_grid("wide61")
    tg = maps.ff_tele_geo(140, 1.32)
    th, phi = maps.ff_fan(100)
    px = maps.ff_pixels(tg, 0.25, 20)
    plan = maps.FarFieldPlan(tg, th, phi, x, y, e, device=0)
    amp, ph, st = plan.run(px)
    assert st == 0
    a1, p1, _ = plan.run(px[:700]); a2, p2, _ = plan.run(px[700:])
    np.testing.assert_array_equal(np.r_[a1, a2], amp); np.testing.assert_array_equal(np.r_[p1, p2], ph)
    d_px = torch.from_numpy(px).cuda(); d_a = torch.empty(len(px), dtype=torch.float64, device="cuda"); d_p = torch.empty_like(d_a)
    plan.run_device(d_px, d_a, d_p, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_a.cpu().numpy(), amp); np.testing.assert_array_equal(d_p.cpu().numpy(), ph)
    assert plan.status() == 0
    sub = px[::40]
    b, q, _ = maps.FarFieldPlan(tg, th, phi, x, y, 3.0 * e * np.exp(0.7j), device=0).run(sub)
    np.testing.assert_allclose(b, 3 * amp[::40], rtol=1e-12)
    assert np.abs(gu.wrap(q - ph[::40] - 0.7)).max() < 1e-12
    # empty receiver field -> exactly zero
    z, _, _ = maps.FarFieldPlan(tg, th, phi, x, y, 0 * e, device=0).run(sub[:3])
    assert np.all(z == 0)


def test_ff_ray_walks_agree(oracle, monkeypatch):
    """The fused map finds the rays that pass the aperture either as one index interval per phi row (monotone
    theta, the reference's linspace fans) or by testing every ray (any other fan).  Both must give the oracle's map:
    on the default fan bit for bit between themselves, on reversed / shuffled / far-off-axis fans against the oracle."""
    from solat_optics_b200 import maps

    tgf = maps.ff_tele_geo(140, 1.32)
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    x, y, e = gu.ff_grid("wide61")
    th, phi = gu.ff_fan(60)
    azel = np.array([[0.3 * gu.DE, gu.EL0], [20 * gu.DE, gu.EL0 - 20 * gu.DE], [-7.3 * gu.DE, gu.EL0 + 11.1 * gu.DE],
                     [40 * gu.DE, gu.EL0 + 40 * gu.DE]])
    a_rows, p_rows, st = maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0).run(azel)
    assert st == 0
    monkeypatch.setenv("SOLAT_FF_QUEUE_WALK", "1")
    a_q, p_q, st = maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0).run(azel)
    monkeypatch.delenv("SOLAT_FF_QUEUE_WALK")
    assert st == 0
    np.testing.assert_array_equal(a_rows, a_q)
    np.testing.assert_array_equal(p_rows, p_q)
    oa, op = oracle.ff_map(geo, azel, th, phi, x, y, e, tgf.k)
    _assert_pixels(a_rows, p_rows, oa, op)
    # the cull of rays beyond the E_rec grid (focal-position predictor) must not change a bit: the +-5 cm grid of
    # ff_sim.py, where half of the accepted rays are culled, with the cull on and off, and against the oracle
    x5, y5, e5 = gu.ff_grid("nf51")
    a_on, p_on, st = maps.FarFieldPlan(tgf, th, phi, x5, y5, e5, device=0).run(azel)
    assert st == 0
    monkeypatch.setenv("SOLAT_FF_PREFILTER", "0")
    a_off, p_off, st = maps.FarFieldPlan(tgf, th, phi, x5, y5, e5, device=0).run(azel)
    monkeypatch.delenv("SOLAT_FF_PREFILTER")
    np.testing.assert_array_equal(a_on, a_off)
    np.testing.assert_array_equal(p_on, p_off)
    oa5, op5 = oracle.ff_map(geo, azel, th, phi, x5, y5, e5, tgf.k)
    _assert_pixels(a_on, p_on, oa5, op5)
    # a fan that crosses the aperture plane's horizon at this pointing: the interval walk must hand the pixel over
    t_h = np.linspace(np.pi - 1.0, np.pi + 1.0, 33)
    a_h, p_h, _ = maps.FarFieldPlan(tgf, t_h, phi, x, y, e, device=0).run([[0.0, 1.2], [0.0, gu.EL0]])
    monkeypatch.setenv("SOLAT_FF_QUEUE_WALK", "1")
    a_hq, p_hq, _ = maps.FarFieldPlan(tgf, t_h, phi, x, y, e, device=0).run([[0.0, 1.2], [0.0, gu.EL0]])
    monkeypatch.delenv("SOLAT_FF_QUEUE_WALK")
    np.testing.assert_array_equal(a_h, a_hq)
    np.testing.assert_array_equal(p_h, p_hq)
    rng = np.random.default_rng(5)
    fans = {"reversed": th[::-1].copy(), "shuffled": rng.permutation(th), "uneven": np.sort(th[0] + (th[-1] - th[0]) * rng.random(60)),
            "wide": np.linspace(np.pi - 9e-5, np.pi + 9e-5, 60), "one_sided": np.linspace(np.pi + 1e-6, np.pi + 5e-5, 60)}
    for name, t in fans.items():
        a, p, st = maps.FarFieldPlan(tgf, t, phi, x, y, e, device=0).run(azel)
        oa, op = oracle.ff_map(geo, azel, t, phi, x, y, e, tgf.k)
        assert st == 0, name
        _assert_pixels(a, p, oa, op)


def test_seeded_sweep_against_oracle(oracle):
    """Seeded random inputs over the ranges the reference's scripts sweep: near-field source and receiver positions,
    frequencies of both bands, odd ray counts; far-field pointings over and beyond the shipped map, focal offsets,
    fan sizes, both E_rec grids.  Every pixel within the north-star tolerance of the oracle."""
    from solat_optics_b200 import maps

    rng = np.random.default_rng(20251020)
    for trial in range(6):
        freq = float(rng.choice([90.0, 140.0, 150.0, 220.0, 280.0]))
        n = int(rng.choice([17, 23, 30, 41]))
        p_rx = tuple(np.r_[rng.uniform(-80, 80), 0.0, rng.uniform(-80, 80)]) if trial % 2 else (0.0, 0.0, 0.0)
        tg = maps.holo_tele_geo(freq, n)
        rx_opl, rx_w = maps.receiver_field(tg, p_rx)
        src = np.c_[rng.uniform(-120, 120, 24), rng.uniform(-120, 120, 24)]
        src[:4] *= 3.0                                              # far off axis: mostly or fully vignetted
        amp, ph, st = maps.NearFieldPlan(tg, [tg.k], rx_opl, rx_w, device=0).run(src)
        assert st == 0
        rx = oracle.nf_rx_trace(p_rx, tg.th_fwhp, n)
        oa, op = oracle.nf_map(src, tg.y_ap, n, rx[3], rx[4] * gu.lyot_mask(n), [tg.k])
        _assert_pixels(amp[0][oa[0] > 0], ph[0][oa[0] > 0], oa[0][oa[0] > 0], op[0][oa[0] > 0])
        assert (amp[0][oa[0] == 0] == 0).all()
    for trial in range(6):
        offset = float(rng.choice([1.25, 1.32, 1.33, 1.4]))
        freq = int(rng.choice([90, 140, 150, 280]))
        n = int(rng.choice([21, 32, 45, 64]))
        tgf = maps.ff_tele_geo(freq, offset)
        geo = oracle.ff_geo_vec(rx_y_m=-offset)
        th, phi = gu.ff_fan(n)
        x, y, e = gu.ff_grid("wide61" if trial % 3 else "nf51")
        azel = np.c_[rng.uniform(-30, 30, 16) * gu.DE, gu.EL0 + rng.uniform(-30, 30, 16) * gu.DE]
        a, p, st = maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0).run(azel)
        assert st == 0
        oa, op = oracle.ff_map(geo, azel, th, phi, x, y, e, tgf.k)
        ok = oa > 0
        _assert_pixels(a[ok], p[ok], oa[ok], op[ok])
        assert (a[~ok] == 0).all()


def test_ff_mirror_pairs(oracle, monkeypatch):
    """The x -> -x symmetry of the LAT lets the fused map trace a ray once for (+az, el) and (-az, el).  Pixels with
    az >= 0 must come out bit-identical with the pair walk on and off, pixels with az < 0 (always computed as mirror
    images) within 1e-9 of the direct computation and within the north-star tolerance of the oracle; a pixel's value
    must not depend on whether its partner is in the list (to the 1e-17 rad the pairing tolerates)."""
    from solat_optics_b200 import maps

    tgf = maps.ff_tele_geo(140, 1.32)
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    th, phi = gu.ff_fan(50)
    x, y, e = gu.ff_grid("nf51")
    px = maps.ff_pixels(tgf, 0.25, 6)                       # 13 x 13, az symmetric about 0 (to ~1e-19 rad)
    a_on, p_on, st = maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0).run(px)
    assert st == 0
    monkeypatch.setenv("SOLAT_FF_PAIRS", "0")
    a_off, p_off, st = maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0).run(px)
    monkeypatch.delenv("SOLAT_FF_PAIRS")
    pos = px[:, 0] >= 0
    np.testing.assert_array_equal(a_on[pos], a_off[pos])
    np.testing.assert_array_equal(p_on[pos], p_off[pos])
    np.testing.assert_allclose(a_on, a_off, rtol=1e-9, atol=0)
    assert np.abs(gu.wrap(p_on - p_off)).max() < 1e-9
    oa, op = oracle.ff_map(geo, px, th, phi, x, y, e, tgf.k)
    _assert_pixels(a_on, p_on, oa, op)
    # partners present or not, list order, duplicates: same values
    plan = maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0)
    neg_only = px[~pos]
    a_n, p_n, _ = plan.run(neg_only)
    np.testing.assert_allclose(a_n, a_on[~pos], rtol=1e-12, atol=0)
    assert np.abs(gu.wrap(p_n - p_on[~pos])).max() < 1e-12
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(px))
    a_p, p_p, _ = plan.run(px[perm])
    np.testing.assert_array_equal(a_p, a_on[perm])
    np.testing.assert_array_equal(p_p, p_on[perm])
    dup = np.concatenate([px[:20], px[:20], px[-5:]])
    a_d, p_d, _ = plan.run(dup)
    np.testing.assert_allclose(a_d, np.concatenate([a_on[:20], a_on[:20], a_on[-5:]]), rtol=1e-12, atol=0)
    # a list above 4096 pixels takes the three-kernel pairing (table in global memory, emptied by the emit phase,
    # alternating counters): repeated calls on one plan, lists of different length in between, same answers
    big = maps.ff_pixels(tgf, 0.1, 33)                      # 67 x 67 = 4489 pixels
    th2, phi2 = gu.ff_fan(24)
    plan2 = maps.FarFieldPlan(tgf, th2, phi2, x, y, e, device=0)
    a1, p1, st = plan2.run(big)
    assert st == 0
    plan2.run(px)
    a2, p2, _ = plan2.run(big)
    a3, p3, _ = plan2.run(big[::-1].copy())
    np.testing.assert_array_equal(a1, a2); np.testing.assert_array_equal(p1, p2)
    np.testing.assert_array_equal(a1, a3[::-1]); np.testing.assert_array_equal(p1, p3[::-1])
    pick = np.r_[0:4489:211, 4488]
    oa, op = oracle.ff_map(geo, big[pick], th2, phi2, x, y, e, tgf.k)
    _assert_pixels(a1[pick], p1[pick], oa, op)
    # (plan2 needs less shared memory than plan: the kernels' dynamic shared-memory limit must not shrink with it)
    # pointings that cannot be paired (az asymmetric beyond the tolerance, huge, NaN) are items of their own
    odd = np.array([[3e-4, gu.EL0], [-3e-4 - 1e-15, gu.EL0], [-5e-4, gu.EL0], [2e2, gu.EL0], [np.nan, gu.EL0]])
    a_o, p_o, _ = plan.run(odd)
    oa, op = oracle.ff_map(geo, odd[:3], th, phi, x, y, e, tgf.k)
    _assert_pixels(a_o[:3], p_o[:3], oa, op)


def test_ff_odd_sizes_and_edge_grids(oracle):
    """Ragged inputs: a fan that is not a multiple of the CTA size, a 1-pixel map, a 2x2 E_rec grid (every ray
    piles into cell 0 or is skipped), a grid entirely below the beam (all rays skipped -> 0)."""
    from solat_optics_b200 import maps

    tg = maps.ff_tele_geo(140, 1.32)
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    th, phi = gu.ff_fan(37)
    # (the 101- and 150-point grids switch the kernel to its 384- and 768-thread CTA shapes: the winner table of a
    #  large grid leaves room for fewer CTAs per SM)
    for ax in (np.array([-1.0, 1.0]), np.linspace(-40, -35, 7), np.linspace(-3, 28, 11), np.linspace(-30, 30, 101),
               np.linspace(-28, 31, 150)):
        yy, xx = np.meshgrid(ax, ax, indexing="ij")
        e = (1 + 0.1 * xx) * np.exp(0.2j * yy)
        a, p, st = maps.FarFieldPlan(tg, th, phi, xx, yy, e, device=0).run([[gu.DE, gu.EL0 - gu.DE]])
        oa, op = oracle.ff_map(geo, [gu.DE, gu.EL0 - gu.DE], th, phi, xx, yy, e, tg.k)
        assert st == 0
        if oa[0] == 0:
            assert a[0] == 0
        else:
            _assert_pixels(a, p, oa, op)


def test_regrid_on_device_matches_fitpack_and_reference():
    """ray_trace_int.regrid with unequal grids runs the device spline; the n_rx=16 -> n_so=30 reference KAT is in
    test_nf_pixel_known_answers (nf_pixels case 12).  Here: layout and agreement with FITPACK."""
    from scipy.interpolate import RectBivariateSpline

    from solat_optics_b200 import ray_trace_int as rti

    rng = np.random.default_rng(3)
    n_rx, n_so = 20, 33
    ax = np.linspace(-70, 70, n_rx)
    X, Z = np.meshgrid(ax, ax)
    rx = np.zeros((20, n_rx * n_rx)); so = rng.normal(size=(26, n_so * n_so))
    rx[3] = (1500 + 0.01 * (X ** 2 + 0.5 * Z ** 2)).ravel(); rx[4] = np.exp(-(X ** 2 + Z ** 2) / 5e3).ravel()
    b1, b2 = rti.regrid(rx, so)
    new = np.linspace(-70, 70, n_so)
    for row in (3, 4):
        ref = RectBivariateSpline(ax, ax, rx[row].reshape(n_rx, n_rx).T, kx=3, ky=3, s=0)(new, new).T
        np.testing.assert_allclose(b2[row].reshape(n_so, n_so), ref, rtol=0, atol=1e-10)
        np.testing.assert_array_equal(b1[row], so[row])


def test_tiny_and_empty_inputs(oracle):
    """Degenerate sizes the reference's loops handle trivially: a 1-ray fan, a 1-pixel map, an empty pixel list,
    a 1x1 E_rec grid, a 3-ray fan in paired mode."""
    from solat_optics_b200 import maps
    from solat_optics_b200 import ray_trace_int as rti

    for n in (1, 2, 3):
        tg = maps.holo_tele_geo(140.0, n)
        out = rti.source_to_lyot_model([2.0, tg.y_ap, -3.0], tg, PlotOpts).output()
        ref = oracle.nf_source_trace([2.0, tg.y_ap, -3.0], n)
        np.testing.assert_allclose(np.nan_to_num(out), np.nan_to_num(ref), rtol=0, atol=RAY_TOL)
        px, amp, ph = maps.holo_map(140.0, res=1 if n == 1 else 2, n_scan=n, device=0)
        rx = oracle.nf_rx_trace([0, 0, 0], tg.th_fwhp, n)
        oa, op = oracle.nf_map(px, tg.y_ap, n, rx[3], rx[4] * gu.lyot_mask(n), [tg.k])
        _assert_pixels(amp, ph, oa, op)
    tg = maps.holo_tele_geo(140.0, 8)
    rx_opl, rx_w = maps.receiver_field(tg)
    plan = maps.NearFieldPlan(tg, [tg.k], rx_opl, rx_w, device=0)
    a, p, st = plan.run(np.zeros((0, 2)))
    assert a.shape == (1, 0) and st == 0
    a, p, st = maps.NearFieldPlan(tg, [tg.k], rx_opl, 0 * rx_w, device=0).run([[0.0, 0.0]])     # nothing passes the mask
    assert a[0, 0] == 0 and p[0, 0] == 0
    # far field
    tgf = maps.ff_tele_geo(140, 1.32)
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    one = np.array([[0.0]])
    # (az = 0 with an odd fan puts rays at focal x = 0 +- 1e-12 mm, exactly on a bin edge of this grid: the
    #  reference's searchsorted is discontinuous there, so the pixel is taken slightly off axis)
    az_t = 0.3 * gu.DE
    for n in (1, 2, 5):
        th, phi = gu.ff_fan(n)
        x, y, e = gu.ff_grid("wide61")
        a, p, st = maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0).run([[az_t, gu.EL0]])
        oa, op = oracle.ff_map(geo, [az_t, gu.EL0], th, phi, x, y, e, tgf.k)
        assert st == 0
        if oa[0] == 0:
            assert a[0] == 0
        else:
            _assert_pixels(a, p, oa, op)
    th, phi = gu.ff_fan(9)
    a, p, st = maps.FarFieldPlan(tgf, th, phi, one + 100.0, one + 100.0, one + 2j, device=0).run([[0.0, gu.EL0], [gu.DE, gu.EL0]])
    oa, op = oracle.ff_map(geo, [[0.0, gu.EL0], [gu.DE, gu.EL0]], th, phi, one + 100.0, one + 100.0, one + 2j, tgf.k)
    _assert_pixels(a, p, oa, op)          # 1x1 grid: every ray below the maximum piles into the single cell
    assert maps.FarFieldPlan(tgf, th, phi, x, y, e, device=0).run(np.zeros((0, 2)))[0].shape == (0,)


# ------------------------------------------------------------------------------------------------ CLI
def test_cli_holo_sim_writes_reference_format(tmp_path):
    d = gu.load("nf_map5")
    out = tmp_path / "nf.txt"
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "holo_sim.py"), "140", "--res", "5", "--n-scan", "20",
                           "--out", str(out)], env=env, cwd=str(tmp_path))
    tab = np.loadtxt(out)
    assert tab.shape == (25, 4)
    np.testing.assert_array_equal(tab[:, 0], d["in_x"]); np.testing.assert_array_equal(tab[:, 1], d["in_z"])
    _assert_pixels(tab[:, 2], tab[:, 3], d["amp"], d["phi"])
    # chain into the far-field CLI (ff_sim.py reads the near-field table as E_rec)
    ff = tmp_path / "ff.txt"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "ff_sim.py"), "140", "0.25", "2", "1.32", "--nf-input", str(out),
                           "--n-scan", "16", "--out", str(ff)], env=env, cwd=str(tmp_path))
    t2 = np.loadtxt(ff)
    assert t2.shape == (25, 4) and np.isfinite(t2).all()
    np.testing.assert_allclose(t2[0, :2], (-2 * gu.DE, gu.EL0 - 2 * gu.DE), atol=1e-15)


# ------------------------------------------------------------------------------------------------ round 2: the timed configurations
def _bench_e_rec():
    """E_rec exactly as bench.py / ff_sim.py build it for config 2: the build's own 51x51 near-field map at 140 GHz
    through near_field_to_e_rec (+-5 cm grid: the predictor cull, the corner-cell rule and the mirror pairs all act)."""
    from solat_optics_b200 import maps

    px, amp, ph = maps.holo_map(140.0, device=0)
    return maps.near_field_to_e_rec(np.c_[px[:, 0], px[:, 1], amp[0], ph[0]])


def test_ff_full_map_on_the_bench_e_rec_against_oracle(oracle):
    """BASELINE config 2 exactly as bench.py times it (ff_sim.py 140 0.25 20 1.32, 41x41 px x 100^2 rays, E_rec = the
    near-field map of config 1 computed by this build): every pixel against the oracle.  far_field_latrt.py:19-30."""
    from solat_optics_b200 import maps

    x, y, e = _bench_e_rec()
    px, amp, ph = maps.ff_map(140, 0.25, 20, 1.32, x, y, e, device=0)
    tg = maps.ff_tele_geo(140, 1.32)
    th, phi = maps.ff_fan(100)
    oa, op = oracle.ff_map(oracle.ff_geo_vec(rx_y_m=-1.32), px, th, phi, x, y, e, tg.k)
    assert np.count_nonzero(oa) == len(oa)
    _assert_pixels(amp, ph, oa, op, max_bad=2)


def _config5():
    import bench
    from solat_optics_b200 import maps

    x, y, e = bench.e_rec_config5()
    tg = maps.ff_tele_geo(bench.FREQ, bench.OFFSET)
    return tg, maps.ff_pixels(tg, bench.RES_ARCMIN, bench.NUM), x, y, e


def test_ff_config5_stratified_pixels_against_oracle(oracle):
    """BASELINE config 5 (401x401 px x 1000^2 rays, 101x101 E_rec grid: the 768-thread CTA shape with two winner tables,
    95 % of the 20-bit ray-index range): 64 stratified pixels -- corners, edges, az = 0, both signs of az -- against
    the oracle at the full fan."""
    import bench
    from solat_optics_b200 import maps

    tg, px, x, y, e = _config5()
    th, phi = maps.ff_fan(bench.N_SCAN)
    idx = bench.stratified_pixel_indices(2 * bench.NUM + 1)
    assert len(idx) == 64 and (2 * bench.NUM + 1) * bench.NUM + bench.NUM in idx
    plan = maps.far_field_plan(tg, th, phi, x, y, e, 0)
    amp, ph, st = plan.run(px[idx])
    assert st == 0
    oa, op = oracle.ff_map(oracle.ff_geo_vec(rx_y_m=-bench.OFFSET), px[idx], th, phi, x, y, e, tg.k)
    _assert_pixels(amp, ph, oa, op)
    # the scout stage (pile rays that cannot win are dropped from the list) must not change a bit
    os.environ["SOLAT_FF_PILE"] = "0"
    try:
        a0, p0, st = maps.FarFieldPlan(tg, th, phi, x, y, e, device=0).run(px[idx[::7]])
    finally:
        del os.environ["SOLAT_FF_PILE"]
    np.testing.assert_array_equal(a0, amp[::7]); np.testing.assert_array_equal(p0, ph[::7])
    # a pixel's value does not depend on the list it comes in: the same pixels inside a longer list, other order
    more = np.r_[px[idx[::-1]], px[1000:1010]]
    a2, p2, st = plan.run(more)
    canon = px[idx[::-1], 0] >= 0
    np.testing.assert_array_equal(a2[:64][canon], amp[::-1][canon])
    np.testing.assert_allclose(a2[:64], amp[::-1], rtol=1e-12)
    assert np.abs(gu.wrap(p2[:64] - ph[::-1])).max() < 1e-12


def test_ff_large_pixel_lists_and_mixed_sizes_on_one_plan(oracle):
    """Lists above 4096 pixels pair on the device with three kernels (ff_pair_insert / attach / emit) instead of one
    CTA, and their per-item records are chunked; calls of both kinds alternate on ONE plan (the small path must not
    leave a stale item counter behind for the large one).  Config-5 pointing grid (121 x 121 of its pixels), 40^2 rays."""
    import bench
    from solat_optics_b200 import maps

    x, y, e = bench.e_rec_config5()
    tg = maps.ff_tele_geo(bench.FREQ, bench.OFFSET)
    th, phi = maps.ff_fan(40)
    px = maps.ff_pixels(tg, bench.RES_ARCMIN * 3, 60)              # 121 x 121 = 14641 pixels
    plan = maps.far_field_plan(tg, th, phi, x, y, e, 0)
    big1 = plan.run(px)
    assert big1[2] == 0
    big2 = plan.run(px[:9001])
    small = plan.run(px[:3000])
    big3 = plan.run(px)
    np.testing.assert_array_equal(big3[0], big1[0]); np.testing.assert_array_equal(big3[1], big1[1])
    canon = px[:, 0] >= 0
    for part, n in ((big2, 9001), (small, 3000)):
        np.testing.assert_array_equal(part[0][canon[:n]], big1[0][:n][canon[:n]])
        np.testing.assert_allclose(part[0], big1[0][:n], rtol=1e-12)
        assert np.abs(gu.wrap(part[1] - big1[1][:n])).max() < 1e-12
    idx = np.arange(0, len(px), 37)
    oa, op = oracle.ff_map(oracle.ff_geo_vec(rx_y_m=-bench.OFFSET), px[idx], th, phi, x, y, e, tg.k)
    ok = oa > 0
    _assert_pixels(big1[0][idx][ok], big1[1][idx][ok], oa[ok], op[ok])


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bench
    from solat_optics_b200 import maps

    maps.SHARD_MIN_PIXEL_RAYS = 0.0                                # deal even these small maps over the ranks
    x, y, e = bench.e_rec_config5()
    ff = maps.ff_map(140, 0.25, 9, 1.32, x, y, e, n_scan=60, device=rank)
    nf = maps.holo_map([150.0, 220.0], res=21, n_scan=40, device=rank)
    if rank == 0:
        q.put((ff, nf))
    else:
        assert ff is None and nf is None
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_sharded_maps_equal_single_rank_maps():
    """the map dealt over 2 ranks (cyclic mirror groups, NCCL gather to rank 0, un-permute) equals the
    1-rank map: bit for bit for az >= 0 / x, z >= 0 (a mirror group is traced whole on one rank either way)."""
    import socket

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import bench
    from solat_optics_b200 import maps

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ff, nf = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    x, y, e = bench.e_rec_config5()
    ff1 = maps.ff_map(140, 0.25, 9, 1.32, x, y, e, n_scan=60, device=0)
    nf1 = maps.holo_map([150.0, 220.0], res=21, n_scan=40, device=0)
    for got, ref in ((ff, ff1), (nf, nf1)):
        np.testing.assert_array_equal(got[0], ref[0])
        np.testing.assert_array_equal(got[1], ref[1])
        np.testing.assert_array_equal(got[2], ref[2])


def _synthetic_layouts():
    """The fixture's synthetic panel layout (gu.panel_identify: ids = floor(coordinate / 1000) + 10) as rectangles."""
    from solat_optics_b200 import solat_apert_field as saf

    ids = [(r, c) for r in range(0, 20) for c in range(0, 20)]
    r = np.array([i[0] for i in ids], dtype=float); c = np.array([i[1] for i in ids], dtype=float)
    lay = saf.PanelLayout((c - 10) * 1000.0, (c - 9) * 1000.0, (r - 10) * 1000.0, (r - 9) * 1000.0, r, c)
    return lay, lay


def test_ff_panel_identification_on_the_device(oracle):
    """Panels identified on the device from a rectangular layout (solat_ff_map_panel_layout_host) against (a) the
    reference fixture ff_panels, (b) the host round trip through an opaque identify_panel, bit for bit, (c) overlapping
    panels (a later panel overrides an earlier one) and a crowded bucket (linear scan), (d) all-zero panel errors with
    missing panels: the cull stays on and only the dropped rays change the map."""
    from solat_optics_b200 import maps
    from solat_optics_b200 import solat_apert_field as saf

    class Opaque:
        @staticmethod
        def get_single_vert_adj_positions():
            z = np.zeros(1)
            return [z, z, z, z, z], [z, z, z, z, z]

        identify_panel = staticmethod(gu.panel_identify)

    d = gu.load("ff_panels")
    pan1, pan2 = d["pan1"], d["pan2"]
    lay1, lay2 = _synthetic_layouts()
    rng = np.random.default_rng(7)
    pts = rng.uniform(-3900, 3900, (2, 50, 50))
    np.testing.assert_array_equal(lay1.identify_panel(pts[0], pts[1]), gu.panel_identify(pts[0], pts[1]))
    n = int(d["in_n_scan"][0])
    th, ph = gu.ff_fan(n)
    x, y, e = gu.ff_grid("wide61")
    tgf = maps.ff_tele_geo(140, 1.32)
    azel = np.c_[d["in_az"], d["in_el"]]
    plan = maps.FarFieldPlan(tgf, th, ph, x, y, e, device=0)
    a, p, st = plan.run_panel_layout(azel, lay1, lay2, pan1, pan2)
    assert st == 0
    _assert_pixels(a, p, d["amp"], d["phi"])                                    # (a) the unmodified reference
    th, ph = gu.ff_fan(64)
    plan = maps.FarFieldPlan(tgf, th, ph, x, y, e, device=0)
    azel = np.array([[0.3 * gu.DE, gu.EL0], [11 * gu.DE, gu.EL0 - 7 * gu.DE], [-20 * gu.DE, gu.EL0 + 20 * gu.DE], [0.0, gu.EL0 + 3 * gu.DE]])
    a_dev, p_dev, st = plan.run_panel_layout(azel, lay1, lay2, pan1, pan2)
    a_host, p_host, st2 = plan.run_panel_models(azel, pan1, pan2, backend=Opaque)
    assert st == 0 and st2 == 0
    np.testing.assert_array_equal(a_dev, a_host); np.testing.assert_array_equal(p_dev, p_host)   # (b)
    # (c) a second, shifted set of panels laid over the first with other ids; 12 thin panels crowd one bucket
    r2 = np.r_[lay1.rect, lay1.rect + [333.0, 333.0, -250.0, -250.0]]
    over = saf.PanelLayout(r2[:, 0], r2[:, 1], r2[:, 2], r2[:, 3], np.r_[lay1.row, lay1.row[::-1]], np.r_[lay1.col, lay1.col[::-1]])
    thin = saf.PanelLayout(np.r_[over.rect[:, 0], np.full(12, -100.0)], np.r_[over.rect[:, 1], 100.0 + np.arange(12)],
                           np.r_[over.rect[:, 2], np.full(12, -100.0)], np.r_[over.rect[:, 3], 100.0 + np.arange(12)],
                           np.r_[over.row, 8 + np.arange(12) % 3], np.r_[over.col, 9 + np.arange(12) % 4])
    for lay in (over, thin):
        class Host(Opaque):
            identify_panel = staticmethod(lambda xx, yy, *a_, _l=lay: _l.identify_panel(xx, yy))

        a_dev, p_dev, st = plan.run_panel_layout(azel, lay, lay, pan1, pan2)
        a_host, p_host, st2 = plan.run_panel_models(azel, pan1, pan2, backend=Host)
        np.testing.assert_array_equal(a_dev, a_host); np.testing.assert_array_equal(p_dev, p_host)
    # (d) zero errors, some panels missing from the model
    z1, z2 = pan1.copy(), pan2.copy()
    z1[2:8] = 0.0; z2[2:8] = 0.0
    a_dev, p_dev, st = plan.run_panel_layout(azel, lay1, lay2, z1, z2)
    a_host, p_host, st2 = plan.run_panel_models(azel, z1, z2, backend=Opaque)
    np.testing.assert_array_equal(a_dev, a_host); np.testing.assert_array_equal(p_dev, p_host)
    plain, _, _ = plan.run(azel)
    assert np.all(np.abs(a_dev - plain) > 1e-6 * plain)


def test_ff_beyond_the_former_limits(oracle):
    """E_rec grids above 168 x 168 (winner tables in global memory instead of shared memory) and fans above
    1 048 575 rays (24-bit ray index in the winner key) used to be errors: both against the oracle, and a grid that
    fits shared memory computed both ways (SOLAT_FF_WIN_GLOBAL=1 forces the global tables)."""
    from solat_optics_b200 import maps

    tg = maps.ff_tele_geo(140, 1.32)
    geo = oracle.ff_geo_vec(rx_y_m=-1.32)
    azel = np.array([[0.0, gu.EL0], [7 * gu.DE, gu.EL0 - 3 * gu.DE], [-7 * gu.DE, gu.EL0 - 3 * gu.DE], [-12 * gu.DE, gu.EL0 + 9 * gu.DE]])
    ax = np.linspace(-30, 30, 201)
    y, x = np.meshgrid(ax, ax, indexing="ij")
    e = np.exp(-(x ** 2 + y ** 2) / 162.0) * np.exp(1j * 0.009 * (x ** 2 + y ** 2))
    th, ph = gu.ff_fan(80)
    a, p, st = maps.FarFieldPlan(tg, th, ph, x, y, e, device=0).run(azel)
    assert st == 0
    oa, op = oracle.ff_map(geo, azel, th, ph, x, y, e, tg.k)
    _assert_pixels(a, p, oa, op)
    # the same map with the tables in shared memory and in global memory
    x6, y6, e6 = gu.ff_grid("wide61")
    a_s, p_s, _ = maps.FarFieldPlan(tg, th, ph, x6, y6, e6, device=0).run(azel)
    os.environ["SOLAT_FF_WIN_GLOBAL"] = "1"
    try:
        a_g, p_g, st = maps.FarFieldPlan(tg, th, ph, x6, y6, e6, device=0).run(azel)
    finally:
        del os.environ["SOLAT_FF_WIN_GLOBAL"]
    assert st == 0
    np.testing.assert_allclose(a_g, a_s, rtol=1e-13)       # (another CTA shape: the winners are summed in another order)
    assert np.abs(gu.wrap(p_g - p_s)).max() < 1e-12
    # 1100 x 1100 = 1.21e6 rays per pixel
    th, ph = gu.ff_fan(1100)
    x5, y5, e5 = gu.ff_grid("nf51")
    a, p, st = maps.FarFieldPlan(tg, th, ph, x5, y5, e5, device=0).run(azel[:3])
    assert st == 0
    oa, op = oracle.ff_map(geo, azel[:3], th, ph, x5, y5, e5, tg.k)
    _assert_pixels(a, p, oa, op)
    # more than 65535 pixels through the mirror-coordinate entry (host identify_panel path)
    th, ph = gu.ff_fan(4)
    plan = maps.FarFieldPlan(tg, th, ph, x6, y6, e6, device=0)
    many = np.c_[np.linspace(-20, 20, 70001) * gu.DE, np.full(70001, gu.EL0)]
    xy, st = plan.mirror_xy(many)
    one, _ = plan.mirror_xy(many[69990:69991])
    np.testing.assert_array_equal(xy[69990], one[0])


def _env_plan(env, *args):
    """A far-field plan created under temporary environment switches (read at plan creation)."""
    from solat_optics_b200 import maps

    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return maps.FarFieldPlan(*args, device=0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def test_ff_interpolated_cells_equal_the_exact_walk(oracle):
    """Large fans take ff_map_interp_kernel: cells from focal positions interpolated on a traced lattice, exact traces
    for the winners only, every winner probed against its prediction, failed items redone by the exact kernel.  The
    maps must not differ by a bit from the exact walk (SOLAT_FF_INTERP=0) -- with the lean binning on and off, with
    every probe forced to fail (all items redone), for single pixels and mirror pairs -- and agree with the oracle."""
    import bench
    from solat_optics_b200 import maps

    tg, px, x, y, e = _config5()
    side = 2 * bench.NUM + 1
    grid = px.reshape(side, side, 2)
    # both signs of az, the az = 0 column, corners, and a list where some pixels have no mirror partner
    sel = np.ascontiguousarray(np.r_[grid[::50, ::100].reshape(-1, 2), grid[3:200:37, 120]])
    for n in (512, 1000):
        th, phi = maps.ff_fan(n)
        pts = sel if n == 512 else sel[::3]
        a_i, p_i, st = _env_plan({}, tg, th, phi, x, y, e).run(pts)
        assert st == 0
        a_x, p_x, st = _env_plan({"SOLAT_FF_INTERP": "0"}, tg, th, phi, x, y, e).run(pts)
        assert st == 0
        np.testing.assert_array_equal(a_i, a_x); np.testing.assert_array_equal(p_i, p_x)
        a_g, p_g, st = _env_plan({"SOLAT_FF_INTERP": "0", "SOLAT_FF_LEAN": "0"}, tg, th, phi, x, y, e).run(pts)
        np.testing.assert_array_equal(a_g, a_x); np.testing.assert_array_equal(p_g, p_x)
        if n == 512:
            a_r, p_r, st = _env_plan({"SOLAT_FF_INTERP_CHK": "0"}, tg, th, phi, x, y, e).run(pts)   # every item redone
            np.testing.assert_array_equal(a_r, a_x); np.testing.assert_array_equal(p_r, p_x)
            a_s, p_s, st = _env_plan({"SOLAT_FF_INTERP": "4"}, tg, th, phi, x, y, e).run(pts)       # finer lattice
            np.testing.assert_array_equal(a_s, a_x); np.testing.assert_array_equal(p_s, p_x)
            oa, op = oracle.ff_map(oracle.ff_geo_vec(rx_y_m=-bench.OFFSET), pts[::4], th, phi, x, y, e, tg.k)
            _assert_pixels(a_i[::4], p_i[::4], oa, op)
    # one pixel per work item (no mirror pairs: one candidate table per CTA)
    th, phi = maps.ff_fan(512)
    a1, p1, st = _env_plan({"SOLAT_FF_PAIRS": "0"}, tg, th, phi, x, y, e).run(sel[::2])
    a0, p0, st0 = _env_plan({"SOLAT_FF_PAIRS": "0", "SOLAT_FF_INTERP": "0"}, tg, th, phi, x, y, e).run(sel[::2])
    assert st == 0 and st0 == 0
    # (one 101 x 101 table per CTA: the exact walk runs 384-thread CTAs here and sums the same winners in another order)
    np.testing.assert_allclose(a1, a0, rtol=1e-14, atol=0); assert np.abs(gu.wrap(p1 - p0)).max() < 1e-14
    # another telescope set-up (90 GHz, focal plane at 1.25 m), a coarser E_rec grid and a pixel list of a different scan
    tg2 = maps.ff_tele_geo(90, 1.25)
    px2 = maps.ff_pixels(tg2, 0.5, 6)
    th, phi = maps.ff_fan(640)
    xc, yc, ec = x[::2, ::2], y[::2, ::2], e[::2, ::2]
    a1, p1, st = _env_plan({}, tg2, th, phi, xc, yc, ec).run(px2[::5])
    a0, p0, st0 = _env_plan({"SOLAT_FF_INTERP": "0"}, tg2, th, phi, xc, yc, ec).run(px2[::5])
    assert st == 0 and st0 == 0
    # (51 x 51 tables: the exact walk runs 256- or 384-thread CTAs here and sums the same winners in another order)
    np.testing.assert_allclose(a1, a0, rtol=1e-14, atol=0); assert np.abs(gu.wrap(p1 - p0)).max() < 1e-14
    # an E_rec grid whose axes are not uniform, and a fan that is not: both keep the exact walk and still agree
    xs = np.sign(x) * np.abs(x) ** 1.1
    th, phi = maps.ff_fan(448)
    a1, p1, st = _env_plan({}, tg, th, phi, xs, y, e).run(sel[:6])
    a0, p0, st = _env_plan({"SOLAT_FF_INTERP": "0", "SOLAT_FF_LEAN": "0"}, tg, th, phi, xs, y, e).run(sel[:6])
    np.testing.assert_array_equal(a1, a0); np.testing.assert_array_equal(p1, p0)
    th2 = th[0] + (th - th[0]) * (1 + 0.02 * (th - th[0]) / (th[-1] - th[0]))
    a1, p1, st = _env_plan({}, tg, th2, phi, x, y, e).run(sel[:6])
    a0, p0, st = _env_plan({"SOLAT_FF_INTERP": "0", "SOLAT_FF_LEAN": "0"}, tg, th2, phi, x, y, e).run(sel[:6])
    np.testing.assert_array_equal(a1, a0); np.testing.assert_array_equal(p1, p0)
    # the lean binning on the small CTA shapes (off by default there: no gain measured) gives the same bits
    th, phi = maps.ff_fan(100)
    xc, yc, ec = x[::2, ::2], y[::2, ::2], e[::2, ::2]
    a1, p1, st = _env_plan({"SOLAT_FF_LEAN": "1"}, tg, th, phi, xc, yc, ec).run(sel)
    a0, p0, st = _env_plan({"SOLAT_FF_LEAN": "0"}, tg, th, phi, xc, yc, ec).run(sel)
    np.testing.assert_array_equal(a1, a0); np.testing.assert_array_equal(p1, p0)
