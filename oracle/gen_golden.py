"""Generate the committed golden fixtures under tests/golden/ by running the
UNMODIFIED reference (/root/reference, via oracle/ref_loader.py shims).

TEST INFRASTRUCTURE ONLY; runs in the build container (the reference cannot travel
to the GPU box).  Usage:  cd /tmp && python /root/repo/oracle/gen_golden.py [--only NAME]

Every fixture stores its inputs next to the reference outputs so the tests need
nothing but the .npz.  Deterministic: the reference has no RNG on this path
(scripts/ff_sim.py:61-62 multiplies its randn by 0).
"""
import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
GOLD = os.path.join(os.path.dirname(_HERE), "tests", "golden")

_ref = None


def _get_ref():
    global _ref
    if _ref is None:
        import warnings

        import ref_loader

        warnings.filterwarnings("ignore")
        _ref = (ref_loader, ref_loader.load())
    return _ref


# ---------------------------------------------------------------- task bodies
def t_nf_rx_rays(a):
    rl, ref = _get_ref()
    tg = rl.nf_tele_geo(ref, a["freq"], a["n_scan"])
    return ref.rti.rx_to_lyot_model(list(a["p_rx"]), tg, rl.PlotOpts).output()


def t_nf_so_rays(a):
    rl, ref = _get_ref()
    tg = rl.nf_tele_geo(ref, a["freq"], a["n_scan"])
    return ref.rti.source_to_lyot_model([a["x"], tg.y_ap, a["z"]], tg, rl.PlotOpts).output()


def t_nf_toy_rays(a):
    rl, ref = _get_ref()
    tg = rl.nf_tele_geo(ref, a["freq"], a["n_scan"])
    return ref.rti.aperature_fields(list(a["p_rx"]), tg, rl.PlotOpts).output()


def t_nf_pixel(a):
    rl, ref = _get_ref()
    return rl.nf_pixel(
        ref, a["freq"], a["x"], a["z"], a["n_scan"], p_rx=a.get("p_rx", (0, 0, 0)),
        n_scan_rx=a.get("n_scan_rx"),
    )


def t_ff_rays(a):
    rl, ref = _get_ref()
    import pan_mod as pm

    tg = rl.ff_tele_geo(ref, a["freq"], a["offset"], a["n_scan"])
    th, ph = rl.ff_fan(a["n_scan"])
    pan2 = pm.panel_model_from_adjuster_offsets(2, np.zeros(1092), 1, 0)
    pan1 = pm.panel_model_from_adjuster_offsets(1, np.zeros(1092), 1, 0)
    m = ref.saf.ray_mirror_pts(tg, th, ph, a["el"], a["az"])
    f = ref.saf.aperature_fields_from_panel_model(pan1, pan2, tg, th, ph, m, a["el"], a["az"])
    return m, f


def ff_grid(kind):
    """Two E_rec grids: 'nf51' mimics what scripts/ff_sim.py:51-59 builds from a default
    51x51 near-field file (+-5 cm; most focal hits fall outside -> exercises the skip /
    pile-into-index-0 behaviour of far_field_latrt.py:19-26); 'wide61' covers +-30 cm so
    the interior binning is exercised.  Field = SURVEY.md §8d config-5 closed form."""
    if kind == "nf51":
        ax = np.linspace(-50, 50, 51) / 1e1
    elif kind == "cfg5":
        ax = np.linspace(-5, 5, 101)                     # bench.e_rec_config5: the E_rec of BASELINE config 5
    else:
        ax = np.linspace(-30, 30, 61)
    # ff_sim.py:55-56: x <- column 1 (fast axis), y <- column 0 (slow axis)
    y, x = np.meshgrid(ax, ax, indexing="ij")
    sig = 1.5 if kind in ("nf51", "cfg5") else 9.0
    amp = np.exp(-(x ** 2 + y ** 2) / (2 * sig ** 2))
    ph = 0.3 * (x ** 2 + y ** 2) * (1.0 if kind in ("nf51", "cfg5") else 0.03)
    e = amp * np.exp(1j * ph)
    return x, y, e


def t_ff_pixel(a):
    rl, ref = _get_ref()
    x, y, e = ff_grid(a["grid"])
    amp, phi, m, f, out_new = rl.ff_pixel(
        ref, a["freq"], a["offset"], a["n_scan"], a["az"], a["el"], x, y, e, want_arrays=True
    )
    return amp, phi, int(np.count_nonzero(f[15])), int(np.count_nonzero(out_new))


def panel_layout(seed=3):
    """Synthetic multi-panel layout for pinning the reference's panel arithmetic (the real layout lives in the
    external pan_mod): 1 m x 1 m panels, ids (row, col) = floor(coordinate / 1000) + 10, a few panels missing
    from the model (rays there are dropped, saf:626-628,805-809), seeded piston / tilt / curvature errors."""
    rng = np.random.default_rng(seed)
    mods = []
    for lo_r, hi_r in ((6, 14), (2, 16)):           # M1: y - y_cent within +-3 m; M2: telescope z of the hit
        ids = [(r, c) for r in range(lo_r, hi_r) for c in range(6, 14) if (r + 2 * c) % 7 != 0]
        pm = np.zeros((10, len(ids)))
        pm[0] = [r for r, _ in ids]; pm[1] = [c for _, c in ids]
        pm[2] = rng.normal(0, 0.02, len(ids))
        pm[3:5] = rng.normal(0, 2e-5, (2, len(ids)))
        pm[5:7] = rng.normal(0, 2e-8, (2, len(ids)))
        pm[7] = rng.normal(0, 1.0, len(ids))        # f: carried by the model, unused by the reference (saf:672)
        pm[8] = (pm[1] - 10 + 0.5) * 1000.0; pm[9] = (pm[0] - 10 + 0.5) * 1000.0
        mods.append(pm)
    return mods[0], mods[1]


def panel_identify(x2d, y2d, x_adj, y_adj, col, row):
    """(2, n, n): [0] = row ids, [1] = col ids (saf:609-617)"""
    return np.stack([np.floor(np.asarray(y2d) / 1000.0) + 10, np.floor(np.asarray(x2d) / 1000.0) + 10])


def t_ff_pixel_panels(a):
    rl, ref = _get_ref()
    x, y, e = ff_grid(a["grid"])
    pan1, pan2 = panel_layout()
    amp, phi, m, f = rl.ff_pixel_panels(ref, a["freq"], a["offset"], a["n_scan"], a["az"], a["el"], x, y, e, pan1, pan2,
                                        panel_identify)
    return amp, phi, m, f


def t_unwrap(a):
    rl, ref = _get_ref()
    return ref.phase_correct.do_unwrap(a["phi_deg"].copy())


def _run(job):
    fn, a = job
    t0 = time.time()
    r = globals()[fn](a)
    return r, time.time() - t0


# ---------------------------------------------------------------- fixture specs
DE = 0.25 / 60 * np.pi / 180  # ff_sim.py:94 with res = 0.25 arcmin
EL0 = float(np.arctan(7.2 / 100e3))


def build_jobs():
    J = {}
    # -- near-field per-ray arrays
    J["nf_rx_rays"] = [
        ("t_nf_rx_rays", dict(freq=140.0, n_scan=12, p_rx=(0, 0, 0))),
        ("t_nf_rx_rays", dict(freq=90.0, n_scan=12, p_rx=(40.0, 0.0, -25.0))),
        ("t_nf_rx_rays", dict(freq=140.0, n_scan=9, p_rx=(-60.0, 3.0, 60.0))),
    ]
    J["nf_so_rays"] = [
        ("t_nf_so_rays", dict(freq=140.0, n_scan=12, x=0.0, z=0.0)),
        ("t_nf_so_rays", dict(freq=140.0, n_scan=12, x=-50.0, z=37.0)),
        ("t_nf_so_rays", dict(freq=140.0, n_scan=12, x=50.0, z=50.0)),
        ("t_nf_so_rays", dict(freq=140.0, n_scan=12, x=300.0, z=-200.0)),
        ("t_nf_so_rays", dict(freq=90.0, n_scan=12, x=12.0, z=-30.0)),
        ("t_nf_so_rays", dict(freq=140.0, n_scan=40, x=-50.0, z=37.0)),
    ]
    J["nf_toy_rays"] = [
        ("t_nf_toy_rays", dict(freq=140.0, n_scan=12, p_rx=(0, 0, 0))),
        ("t_nf_toy_rays", dict(freq=90.0, n_scan=10, p_rx=(25.0, 0.0, -10.0))),
    ]
    # -- near-field per-pixel known answers (SURVEY.md §8c list + extras)
    px = []
    for x, z in [(0, 0), (-50, 37)]:
        px.append(dict(freq=140.0, n_scan=100, x=float(x), z=float(z)))
    for x, z in [(0, 0), (-50, 37), (50, 50)]:
        px.append(dict(freq=140.0, n_scan=40, x=float(x), z=float(z)))
    for x, z in [(0, 0), (-50, -50), (12, -30)]:
        px.append(dict(freq=140.0, n_scan=30, x=float(x), z=float(z)))
    for f in [90.0, 150.0, 220.0, 280.0]:
        px.append(dict(freq=f, n_scan=30, x=12.0, z=-30.0))
    px.append(dict(freq=140.0, n_scan=30, x=12.0, z=-30.0, n_scan_rx=16))  # unequal grids
    px.append(dict(freq=150.0, n_scan=30, x=-20.0, z=44.0, p_rx=(60.0, 0.0, -30.0)))  # rx offset
    px.append(dict(freq=150.0, n_scan=30, x=600.0, z=0.0))  # fully vignetted
    px.append(dict(freq=140.0, n_scan=30, x=250.0, z=-100.0))  # mostly vignetted
    J["nf_pixels"] = [("t_nf_pixel", a) for a in px]
    # -- receiver-offset sweep (quant_systematics, BASELINE config 4)
    J["nf_rx_offsets"] = [
        ("t_nf_pixel", dict(freq=150.0, n_scan=30, x=float(x), z=float(z), p_rx=p))
        for p in [(-60.0, 0.0, -60.0), (60.0, 0.0, 60.0), (0.0, 0.0, 60.0), (100.0, 0.0, 0.0)]
        for x, z in [(0, 0), (-37, 21)]
    ]
    # -- near-field small map, holo_sim.py grid logic with RES=5, n_scan=20
    rx = np.linspace(-50, 50, 5)
    J["nf_map5"] = [
        ("t_nf_pixel", dict(freq=140.0, n_scan=20, x=float(x), z=float(z))) for x in rx for z in rx
    ]
    # -- far-field per-ray arrays
    J["ff_rays"] = [
        ("t_ff_rays", dict(freq=140, offset=1.32, n_scan=12, az=0.0, el=EL0)),
        ("t_ff_rays", dict(freq=140, offset=1.32, n_scan=12, az=20 * DE, el=EL0 - 20 * DE)),
        ("t_ff_rays", dict(freq=140, offset=1.32, n_scan=12, az=-3 * DE, el=EL0 + 7 * DE)),
        ("t_ff_rays", dict(freq=90, offset=1.25, n_scan=11, az=5 * DE, el=EL0 + 2 * DE)),
    ]
    # -- far-field per-pixel known answers
    fp = []
    for grid in ["nf51", "wide61"]:
        for az, el in [(0.0, EL0), (20 * DE, EL0 - 20 * DE), (-3 * DE, EL0 + 7 * DE), (DE, EL0)]:
            fp.append(dict(freq=140, offset=1.32, n_scan=24, az=az, el=el, grid=grid))
    fp.append(dict(freq=140, offset=1.32, n_scan=100, az=0.0, el=EL0, grid="nf51"))
    fp.append(dict(freq=140, offset=1.32, n_scan=100, az=-7 * DE, el=EL0 + 4 * DE, grid="wide61"))
    fp.append(dict(freq=90, offset=1.25, n_scan=24, az=2 * DE, el=EL0 - DE, grid="wide61"))
    J["ff_pixels"] = [("t_ff_pixel", a) for a in fp]
    # -- the tutorial's command line `ff_sim.py 150 1 10 1.33` (latrt_farfield_sim.ipynb): centre, edge, corners
    D1 = 1.0 / 60 * np.pi / 180
    J["ff_pixels_150"] = [
        ("t_ff_pixel", dict(freq=150, offset=1.33, n_scan=30, az=a * D1, el=EL0 + e * D1, grid="wide61"))
        for a, e in [(0, 0), (10, 0), (-10, 10), (10, -10), (3, -7)]
    ]
    # -- the large-fan path (fans of 400^2 rays and more take the interpolated-cell kernel): three pixels of the
    #    401 x 401 config-5 scan -- a mirror pair and an az = 0 pixel -- at 400^2 rays on the config-5 E_rec
    D5 = 0.05 / 60 * np.pi / 180
    J["ff_pixels_big"] = [
        ("t_ff_pixel", dict(freq=140, offset=1.32, n_scan=400, az=a * D5, el=EL0 + e * D5, grid="cfg5"))
        for a, e in [(-190, -163), (190, -163), (0, 50)]
    ]
    # -- far-field small map: ff_sim.py 140 0.25 2 1.32 with N_scan=16, wide grid
    azs = np.linspace(-2 * DE, 2 * DE, 5)
    els = np.linspace(-2 * DE + EL0, 2 * DE + EL0, 5)
    J["ff_map5"] = [
        ("t_ff_pixel", dict(freq=140, offset=1.32, n_scan=16, az=float(a), el=float(e), grid="wide61"))
        for a in azs
        for e in els
    ]
    # -- far-field pixels with non-zero panel errors on a synthetic multi-panel layout (rays + pixel answers)
    J["ff_panels"] = [
        ("t_ff_pixel_panels", dict(freq=140, offset=1.32, n_scan=16, az=float(az), el=float(el), grid="wide61"))
        for az, el in [(0.3 * DE, EL0), (20 * DE, EL0 - 20 * DE), (-3 * DE, EL0 + 7 * DE)]
    ]
    # -- phase unwrap (phase_correct.py:8-92): smooth quadratic phase wrapped to +-180 deg
    ax = np.linspace(-1, 1, 21)
    xx, yy = np.meshgrid(ax, ax)
    true = 900.0 * (xx ** 2 + 0.7 * yy ** 2) + 40 * xx
    J["unwrap"] = [("t_unwrap", dict(phi_deg=(np.mod(true + 180, 360) - 180)))]
    return J


def pack(name, jobs, results):
    out = {}
    args = [a for _, a in jobs]
    keys = sorted({k for a in args for k in a if k != "phi_deg"})
    for k in keys:
        vals = [a.get(k, None) for a in args]
        if k == "p_rx":
            vals = [(0, 0, 0) if v is None else v for v in vals]
        if k == "n_scan_rx":
            vals = [a["n_scan"] if v is None else v for v, a in zip(vals, args)]
        out["in_" + k] = np.array(vals)
    if name in ("nf_rx_rays", "nf_so_rays", "nf_toy_rays"):
        for i, r in enumerate(results):
            out["out_%d" % i] = r
    elif name in ("nf_pixels", "nf_map5", "nf_rx_offsets"):
        out["amp"] = np.array([r[0] for r in results])
        out["phi"] = np.array([r[1] for r in results])
    elif name == "ff_rays":
        for i, (m, f) in enumerate(results):
            out["mirror_%d" % i] = m
            out["field_%d" % i] = f
    elif name in ("ff_pixels", "ff_map5", "ff_pixels_150", "ff_pixels_big"):
        out["amp"] = np.array([r[0] for r in results])
        out["phi"] = np.array([r[1] for r in results])
        out["n_valid"] = np.array([r[2] for r in results])
        out["n_cells"] = np.array([r[3] for r in results])
    elif name == "ff_panels":
        out["amp"] = np.array([r[0] for r in results])
        out["phi"] = np.array([r[1] for r in results])
        for i, r in enumerate(results):
            out["mirror_%d" % i] = r[2]
            out["field_%d" % i] = r[3]
        out["pan1"], out["pan2"] = panel_layout()
    elif name == "unwrap":
        out["phi_deg"] = jobs[0][1]["phi_deg"]
        out["unwrapped_deg"] = results[0]
    out["n_cases"] = np.array(len(jobs))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    J = build_jobs()
    names = [a.only] if a.only else list(J)
    flat = [(n, i, J[n][i]) for n in names for i in range(len(J[n]))]
    # longest first
    flat.sort(key=lambda t: -(t[2][1].get("n_scan", 1) ** 2))
    t0 = time.time()
    with mp.get_context("fork").Pool(a.procs) as pool:
        res = pool.map(_run, [j for _, _, j in flat], chunksize=1)
    got = {n: [None] * len(J[n]) for n in names}
    cpu = 0.0
    for (n, i, _), (r, dt) in zip(flat, res):
        got[n][i] = r
        cpu += dt
    for n in names:
        pack(n, J[n], got[n])
        print("wrote", n, len(J[n]), "cases")
    print("wall %.1f s, reference cpu %.1f s" % (time.time() - t0, cpu))


if __name__ == "__main__":
    main()
