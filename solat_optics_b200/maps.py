"""Whole-map plans and multi-GPU sharding.

Replaces the reference's MPI master/worker fan-out (scripts/holo_sim.py:87-189,
scripts/ff_sim.py:135-227, scripts/beamsim_set.sh): the scan-pixel list is cut into contiguous
blocks, one per rank (one process per GPU under torchrun), every rank runs ONE fused kernel over
its block, and a single gather (NCCL over NVLink; gloo in the CPU tests) brings (amp, phase) to
rank 0.  Pixels are independent, so there is no other exchange step.
"""
import ctypes
import os

import numpy as np

from . import _lib
from . import ray_trace_int as rti

R_LYOT = 136.54 / 2  # scripts/holo_sim.py:168


class PlotOpts:
    """Same role as the class in scripts/holo_sim.py:25-31 (no plotting)."""

    plot = 0
    color = "darkorange"
    alpha = 0.2


# ------------------------------------------------------------------------------------------------
# sharding / gather
# ------------------------------------------------------------------------------------------------
def shard_bounds(n_items, rank, world_size):
    """Contiguous block [lo, hi) of the flattened pixel index owned by `rank`."""
    blk = -(-n_items // world_size)
    lo = min(rank * blk, n_items)
    return lo, min(lo + blk, n_items)


def dist_info():
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun env, else (0, 1)."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def gather_blocks(local, n_items, dst=0):
    """Gather per-rank blocks (torch tensors [n_local, C], block layout of shard_bounds) to rank `dst`.
    Returns the [n_items, C] tensor on `dst`, None elsewhere.  One collective."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local[:n_items]
    world, rank = dist.get_world_size(), dist.get_rank()
    blk = -(-n_items // world)
    pad = torch.zeros((blk, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat(out, dim=0)[:n_items]


# ------------------------------------------------------------------------------------------------
# near field
# ------------------------------------------------------------------------------------------------
def holo_tele_geo(freq, n_scan=100):
    """tele_geo exactly as scripts/holo_sim.py:34-57 sets it up for `freq` [GHz]."""
    from . import ot_geo

    tg = ot_geo.LatGeo()
    if freq <= 120:
        tg.y_ap = -1213 - (6.5 * 10)
        tg.th_fwhp = 40.5 * np.pi / 180
    else:
        tg.y_ap = -1213 - (12.2 * 10)
        tg.th_fwhp = 25 * np.pi / 180
    tg.x_ap = 0
    tg.z_ap = 0
    tg.lambda_ = 0.299792458 / freq
    tg.k = 2 * np.pi / tg.lambda_
    tg.n_scan = n_scan
    return tg


def receiver_field(tele_geo, p_rx=(0, 0, 0), n_scan_rx=None, n_scan_so=None):
    """(rx_opl, rx_w) in the source's ray-index space: the pixel-independent half of
    scripts/holo_sim.py:157-180 (rx trace -> regrid -> Lyot mask x feed taper)."""
    n_so = int(tele_geo.n_scan if n_scan_so is None else n_scan_so)
    n_rx = int(n_so if n_scan_rx is None else n_scan_rx)
    keep = tele_geo.n_scan
    try:
        tele_geo.n_scan = n_rx
        out_rx = rti.rx_to_lyot_model(list(p_rx), tele_geo, PlotOpts).output()
    finally:
        tele_geo.n_scan = keep
    if n_rx == n_so:
        # equal grids: regrid is the identity on rows 3, 4 and puts the nominal +-70 mm index grid in rows 0, 2
        return np.ascontiguousarray(out_rx[3, :]), np.where(_lyot_mask(n_so), out_rx[4, :], 0.0)
    so_like = np.zeros((26, n_so * n_so))
    _, rx_new = rti.regrid(out_rx, so_like)   # second return value carries the rx field (rti:1435-1438)
    lyot_r = np.sqrt(rx_new[0, :] ** 2 + rx_new[2, :] ** 2)
    w = np.where(lyot_r <= R_LYOT, rx_new[4, :], 0.0)
    return np.ascontiguousarray(rx_new[3, :]), np.ascontiguousarray(w)


def receiver_fields(tele_geo, rx_positions, n_scan=None, device=None):
    """receiver_field for many receiver positions in one launch (equal receiver and source grids):
    (rx_opl [n_rx, n_scan^2], rx_w [n_rx, n_scan^2])."""
    n = int(tele_geo.n_scan if n_scan is None else n_scan)
    pos = _lib.as_f64(np.reshape(rx_positions, (-1, 3)))
    # launch fan of rx_to_lyot_model (rti:593-598)
    theta = np.linspace(-(np.pi / 2) - 0.28, -(np.pi / 2) + 0.28, n)
    phi = np.linspace((np.pi / 2) - 0.28, (np.pi / 2) + 0.28, n)
    opl = np.empty((len(pos), n * n))
    amp = np.empty((len(pos), n * n))
    st = _lib.c_i(0)
    geom = _lib.nf_geom(tele_geo)
    _lib.check(_lib.load().solat_nf_rx_fields(_lib.context(device).handle, ctypes.byref(geom), _lib.dptr(pos), len(pos),
                                             _lib.dptr(theta), n, _lib.dptr(phi), n, _lib.dptr(opl), _lib.dptr(amp), ctypes.byref(st)))
    _lib.raise_on_status(st.value, rti.strict)
    return opl, np.where(_lyot_mask(n)[None, :], amp, 0.0)


_LYOT_MASKS = {}


def _lyot_mask(n):
    """sqrt(X^2 + Z^2) <= R_LYOT on regrid's nominal index grid (rti:1403-1415, holo_sim.py:168-171)."""
    m = _LYOT_MASKS.get(n)
    if m is None:
        axis = np.linspace(-70, 70, n)
        X2, Z2 = np.meshgrid(axis, axis)
        m = _LYOT_MASKS[n] = np.sqrt(np.ravel(X2) ** 2 + np.ravel(Z2) ** 2) <= R_LYOT
    return m


class NearFieldPlan:
    """Fused near-field map: per source position sum_rays w_rx exp(i k (L_rx + L_source)/1e3).

    `k_per_m`: one or several wavenumbers [rad/m] sharing the geometry (same y_ap).
    `rx_opl`, `rx_w`: receiver field in the source's ray-index space, shape (n_scan^2,) or, for several
    receiver positions sharing the same source trace, (n_rx, n_scan^2).
    Results have shape (n_freq, npix), or (n_rx, n_freq, npix) when several receiver tables were given."""

    def __init__(self, tele_geo, k_per_m, rx_opl, rx_w, n_scan=None, device=None):
        lib = _lib.load()
        self.ctx = _lib.context(device)
        n = int(tele_geo.n_scan if n_scan is None else n_scan)
        self.n_scan = n
        axis = np.linspace((np.pi / 2) - 0.22, (np.pi / 2) + 0.22, n)   # rti:908-909
        self.k = _lib.as_f64(np.atleast_1d(k_per_m))
        rx_opl = _lib.as_f64(rx_opl)
        rx_w = _lib.as_f64(rx_w)
        self.multi = rx_opl.ndim == 2
        rx_opl = np.atleast_2d(rx_opl)
        rx_w = np.atleast_2d(rx_w)
        self.n_rx = rx_opl.shape[0]
        if rx_opl.shape != (self.n_rx, n * n) or rx_w.shape != rx_opl.shape:
            raise ValueError("receiver field must have n_scan^2 entries (per receiver table)")
        geom = _lib.nf_geom(tele_geo)
        h = _lib.c_vp()
        _lib.check(lib.solat_nf_plan_create_multi(self.ctx.handle, ctypes.byref(geom), _lib.dptr(axis), n, _lib.dptr(axis),
                                                  n, _lib.dptr(rx_opl), _lib.dptr(rx_w), self.n_rx, _lib.dptr(self.k),
                                                  self.k.size, ctypes.byref(h)))
        self.handle = h
        self.active_rays = int(lib.solat_nf_plan_active_rays(h))
        self.channels = int(lib.solat_nf_plan_channels(h))

    def _shape(self, npix):
        return (self.n_rx, self.k.size, npix) if self.multi else (self.k.size, npix)

    def run(self, src_xz, amp=None, phase=None):
        """Host buffers in, host buffers out (copies inside).  Returns amp, phase, status."""
        src = _lib.as_f64(np.reshape(src_xz, (-1, 2)))
        npix = src.shape[0]
        amp = np.empty(self._shape(npix)) if amp is None else amp
        phase = np.empty(self._shape(npix)) if phase is None else phase
        if npix == 0:
            return amp, phase, 0
        st = _lib.c_i(0)
        _lib.check(_lib.load().solat_nf_map_host(self.handle, _lib.dptr(src), npix, _lib.dptr(amp), _lib.dptr(phase),
                                                 ctypes.byref(st)))
        return amp, phase, st.value

    def run_device(self, d_src, d_amp, d_phase, stream=0):
        """torch CUDA float64 tensors ([npix,2], [channels*npix], [channels*npix]); asynchronous on `stream`."""
        npix = d_src.shape[0]
        _lib.check(_lib.load().solat_nf_map_device(self.handle, d_src.data_ptr(), npix, d_amp.data_ptr(),
                                                   d_phase.data_ptr(), stream))

    def status(self):
        st = _lib.c_i(0)
        _lib.check(_lib.load().solat_nf_plan_status(self.handle, ctypes.byref(st)))
        return st.value

    def close(self):
        if self.handle:
            _lib.load().solat_nf_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def holo_pixels(res=51, sweep=100, shift=0):
    """Source-position grid of scripts/holo_sim.py:90-108: row index i*res+j <-> (x_i, z_j)."""
    rx = np.linspace(-sweep / 2 + shift, sweep / 2 + shift, res)
    xx, zz = np.meshgrid(rx, rx, indexing="ij")
    return np.stack([xx.ravel(), zz.ravel()], axis=1)


def holo_map(freqs, res=51, sweep=100, shift=0, n_scan=100, n_scan_rx=None, p_rx=(0, 0, 0), device=None,
             strict=False):
    """Near-field holography map(s) for one band.  `freqs`: scalar or list of GHz values that share
    the band geometry (holo_sim.py:40-45).  Shards over ranks when torch.distributed is initialised.
    Returns on rank 0: (pixels [npix,2], amp [nfreq,npix], phase [nfreq,npix]); None elsewhere."""
    freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    geos = [holo_tele_geo(f, n_scan) for f in freqs]
    if len({g.y_ap for g in geos}) != 1:
        raise ValueError("frequencies of different bands (y_ap differs) need separate maps")
    tg = geos[0]
    # the feed taper differs per band only; one geometry -> one rx field
    rx_opl, rx_w = receiver_field(tg, p_rx, n_scan_rx, n_scan)
    plan = NearFieldPlan(tg, [g.k for g in geos], rx_opl, rx_w, n_scan, device)
    pixels = holo_pixels(res, sweep, shift)
    rank, world = dist_info()
    # ranks take contiguous blocks of the list ordered so that the mirror positions (+-x, +-z) are neighbours: the
    # kernel traces such a group once
    order = mirror_pair_order(pixels, fold_second=True) if world > 1 else np.arange(len(pixels))
    lo, hi = shard_bounds(len(pixels), rank, world)
    if hi > lo:
        amp, phase, st = plan.run(pixels[order[lo:hi]])
    else:
        amp, phase, st = np.zeros((len(freqs), 0)), np.zeros((len(freqs), 0)), 0
    _lib.raise_on_status(st, strict)
    plan.close()
    if world == 1:
        return pixels, amp, phase
    import torch

    dev = _gather_device(plan.ctx.device)
    local = torch.from_numpy(np.concatenate([amp.T, phase.T], axis=1)).to(dev)
    full = gather_blocks(local, len(pixels))
    if full is None:
        return None
    full = full.cpu().numpy()
    out = np.empty_like(full)
    out[order] = full
    nf = len(freqs)
    return pixels, np.ascontiguousarray(out[:, :nf].T), np.ascontiguousarray(out[:, nf:].T)


def receiver_offset_sweep(freq, rx_positions, res=51, sweep=100, shift=0, n_scan=100, device=None, strict=False):
    """Near-field maps for many receiver feed positions at one frequency (the quant_systematics sweep,
    BASELINE config 4): the source rays are traced once per pixel and integrated against every receiver
    table.  Single process (shard the receiver list over ranks for more GPUs).
    Returns (pixels [npix,2], amp [n_rx,npix], phase [n_rx,npix])."""
    tg = holo_tele_geo(float(freq), n_scan)
    rx_opl, rx_w = receiver_fields(tg, rx_positions, n_scan, device)
    plan = NearFieldPlan(tg, [tg.k], rx_opl, rx_w, n_scan, device)
    pixels = holo_pixels(res, sweep, shift)
    amp, phase, st = plan.run(pixels)
    _lib.raise_on_status(st, strict)
    plan.close()
    return pixels, amp[:, 0, :], phase[:, 0, :]


def _gather_device(device_index):
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", device_index)
    return torch.device("cpu")


# ------------------------------------------------------------------------------------------------
# far field
# ------------------------------------------------------------------------------------------------
def ff_tele_geo(freq, offset):
    """tele_geo as scripts/ff_sim.py:29-44,72-79 sets it up."""
    from . import latrt_geo

    tg = latrt_geo.initialize_telescope_geometry()
    tg.rx_x = 0
    tg.rx_z = 0
    tg.rx_y = 0 - offset
    tg.lambda_ = (30.0 / freq) * 0.01
    tg.k = 2 * np.pi / tg.lambda_
    tg.N_scan = 100
    tg.x_tow, tg.y_tow, tg.z_tow = 0, -7.2, 100e3
    tg.el0 = np.arctan(-tg.y_tow / tg.z_tow)
    return tg


def ff_fan(n_scan=100):
    """Polar launch fan of scripts/ff_sim.py:64-66."""
    return np.linspace(np.pi - 0.00005, np.pi + 0.00005, n_scan), np.linspace(0, np.pi, n_scan)


def ff_pixels(tele_geo, res_arcmin, num):
    """(az, el) grid of scripts/ff_sim.py:94,121-128: row index i*len+j <-> (az_i, el_j)."""
    de_ang = res_arcmin / 60 * np.pi / 180
    n = int(num)
    azs = np.linspace(-n * de_ang + tele_geo.az0, n * de_ang + tele_geo.az0, 2 * n + 1)
    els = np.linspace(-n * de_ang + tele_geo.el0, n * de_ang + tele_geo.el0, 2 * n + 1)
    aa, ee = np.meshgrid(azs, els, indexing="ij")
    return np.stack([aa.ravel(), ee.ravel()], axis=1)


def near_field_to_e_rec(foc_fields):
    """Turn a near-field .txt table into the (data_x, data_y, E_rec) of scripts/ff_sim.py:51-59,81-88."""
    from . import phase_correct

    foc_fields = np.asarray(foc_fields, dtype=np.float64)
    n = int(np.sqrt(len(foc_fields)))
    x = np.reshape(foc_fields[:, 1], (n, n)) / 1e1
    y = np.reshape(foc_fields[:, 0], (n, n)) / 1e1
    amp = np.reshape(foc_fields[:, 2], (n, n))
    phi = np.reshape(foc_fields[:, 3], (n, n))
    pnew = phase_correct.do_unwrap(phi * 180 / np.pi) * np.pi / 180
    e_rec = abs(amp) * np.exp(complex(0, 1) * np.mod(pnew, 2 * np.pi))
    return x, y, e_rec


class FarFieldPlan:
    """Fused far-field map: per (az, el) the winners-only sum of ff_sim.py:208-221."""

    def __init__(self, tele_geo, theta, phi, data_x, data_y, e_rec, device=None):
        lib = _lib.load()
        self.ctx = _lib.context(device)
        theta = _lib.as_f64(theta)
        phi = _lib.as_f64(phi)
        data_x = _lib.as_f64(np.real(data_x))
        data_y = _lib.as_f64(np.real(data_y))
        e = np.ascontiguousarray(e_rec, dtype=np.complex128)
        m = data_x.shape[0]
        if data_x.shape != (m, m) or data_y.shape != (m, m) or e.shape != (m, m):
            raise ValueError("data_x, data_y, e_rec must be (M, M)")
        self.n_rays = theta.size * phi.size
        geom = _lib.ff_geom(tele_geo)
        h = _lib.c_vp()
        _lib.check(lib.solat_ff_plan_create(self.ctx.handle, ctypes.byref(geom), _lib.dptr(theta), theta.size,
                                            _lib.dptr(phi), phi.size, _lib.dptr(data_x), _lib.dptr(data_y),
                                            e.ctypes.data_as(_lib.c_dp), m, float(tele_geo.k), ctypes.byref(h)))
        self.handle = h

    def run(self, azel, amp=None, phase=None):
        azel = _lib.as_f64(np.reshape(azel, (-1, 2)))
        npix = azel.shape[0]
        amp = np.empty(npix) if amp is None else amp
        phase = np.empty(npix) if phase is None else phase
        if npix == 0:
            return amp, phase, 0
        st = _lib.c_i(0)
        _lib.check(_lib.load().solat_ff_map_host(self.handle, _lib.dptr(azel), npix, _lib.dptr(amp), _lib.dptr(phase),
                                                 ctypes.byref(st)))
        return amp, phase, st.value

    def run_device(self, d_azel, d_amp, d_phase, stream=0):
        _lib.check(_lib.load().solat_ff_map_device(self.handle, d_azel.data_ptr(), d_azel.shape[0], d_amp.data_ptr(),
                                                   d_phase.data_ptr(), stream))

    def mirror_xy(self, azel):
        """(npix, 4, n_rays): M2 hit x, z and M1 hit x, y of the error-free trace -- the coordinates the reference
        hands to pan_mod.identify_panel (saf:596-612).  Returns (xy, status)."""
        azel = _lib.as_f64(np.reshape(azel, (-1, 2)))
        xy = np.empty((azel.shape[0], 4, self.n_rays))
        st = _lib.c_i(0)
        if azel.shape[0]:
            _lib.check(_lib.load().solat_ff_map_mirror_xy_host(self.handle, _lib.dptr(azel), azel.shape[0], _lib.dptr(xy),
                                                               ctypes.byref(st)))
        return xy, st.value

    def run_panels(self, azel, pan_idx1, pan_idx2, panel_model1, panel_model2):
        """Fused map with panel models: pan_idx1/2 (npix, n_rays) = column of panel_model1/2 every ray sits on
        (-1: none, the ray is dropped); the panel-error polynomials enter both mirror intersections."""
        azel = _lib.as_f64(np.reshape(azel, (-1, 2)))
        npix = azel.shape[0]
        amp, phase = np.empty(npix), np.empty(npix)
        if npix == 0:
            return amp, phase, 0
        i1 = np.ascontiguousarray(pan_idx1, dtype=np.int16).reshape(npix, self.n_rays)
        i2 = np.ascontiguousarray(pan_idx2, dtype=np.int16).reshape(npix, self.n_rays)
        pm1, pm2 = _lib.as_f64(panel_model1), _lib.as_f64(panel_model2)
        if pm1.ndim != 2 or pm2.ndim != 2 or pm1.shape[0] != 10 or pm2.shape[0] != 10:
            raise ValueError("panel models must be (10, n_panels)")
        st = _lib.c_i(0)
        _lib.check(_lib.load().solat_ff_map_panels_host(
            self.handle, _lib.dptr(azel), npix, i1.ctypes.data_as(_lib.c_sp), i2.ctypes.data_as(_lib.c_sp), _lib.dptr(pm1),
            pm1.shape[1], _lib.dptr(pm2), pm2.shape[1], _lib.dptr(amp), _lib.dptr(phase), ctypes.byref(st)))
        return amp, phase, st.value

    def run_panel_models(self, azel, panel_model1, panel_model2, backend=None, batch=64):
        """ff_sim.py:204-221 for a list of pixels with the panel models of pan_mod: mirror coordinates on the GPU ->
        `identify_panel` of the panel backend on the host, pixel by pixel exactly as the reference calls it
        (saf:587-621) -> fused GPU map.  `batch` pixels share one round trip."""
        from . import solat_apert_field as saf

        backend = backend or saf.panel_backend()
        azel = _lib.as_f64(np.reshape(azel, (-1, 2)))
        adj1, adj2 = backend.get_single_vert_adj_positions()
        n = int(round(np.sqrt(self.n_rays)))
        if n * n != self.n_rays:
            raise ValueError("panel identification needs a square fan (saf:596-607)")
        amp, phase, status = np.empty(len(azel)), np.empty(len(azel)), 0
        for lo in range(0, len(azel), batch):
            px = azel[lo:lo + batch]
            xy, st = self.mirror_xy(px)
            status |= st
            i1 = np.empty((len(px), self.n_rays), dtype=np.int16)
            i2 = np.empty_like(i1)
            for p in range(len(px)):
                id2 = backend.identify_panel(xy[p, 0].reshape(n, n), xy[p, 1].reshape(n, n), adj2[4], adj2[3], adj2[0], adj2[1])
                id1 = backend.identify_panel(xy[p, 2].reshape(n, n), xy[p, 3].reshape(n, n) - saf.y_cent_m1, adj1[2], adj1[3],
                                             adj1[0], adj1[1])
                i1[p] = saf._first_match_index(panel_model1, np.ravel(id1[0]), np.ravel(id1[1]))
                i2[p] = saf._first_match_index(panel_model2, np.ravel(id2[0]), np.ravel(id2[1]))
            a, ph, st = self.run_panels(px, i1, i2, panel_model1, panel_model2)
            status |= st
            amp[lo:lo + batch], phase[lo:lo + batch] = a, ph
        return amp, phase, status

    def status(self):
        st = _lib.c_i(0)
        _lib.check(_lib.load().solat_ff_plan_status(self.handle, ctypes.byref(st)))
        return st.value

    def close(self):
        if self.handle:
            _lib.load().solat_ff_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mirror_pair_order(pixels, fold_second=False):
    """Permutation of an (az, el) -- or near-field (x, z) -- pixel list that puts (+az, el) next to (-az, el)
    (|az| compared to 1e-9 of the map's extent, np.linspace grids are not symmetric to the bit).  fold_second: the
    four near-field positions (+-x, +-z) next to each other."""
    pixels = np.asarray(pixels, dtype=np.float64)

    def folded(col):
        scale = float(np.max(np.abs(col))) if len(col) else 0.0
        return np.round(np.abs(col) * (1e9 / scale if scale > 0 else 1.0))

    if fold_second:
        return np.lexsort((np.signbit(pixels[:, 0]), np.signbit(pixels[:, 1]), folded(pixels[:, 1]), folded(pixels[:, 0])))
    return np.lexsort((np.signbit(pixels[:, 0]), pixels[:, 1], folded(pixels[:, 0])))


def ff_map(freq, res_arcmin, num, offset, data_x, data_y, e_rec, n_scan=100, device=None, strict=False, panel_models=None):
    """Far-field beam map `ff_sim.py <freq> <res> <num> <offset>`.  Returns on rank 0
    (pixels [npix,2] (az, el), amp [npix], phase [npix]); None on the other ranks.
    `panel_models` = (pan_mod1, pan_mod2) of ff_sim.py:69-70 switches on panel identification through the panel
    backend (solat_apert_field.set_panel_backend / an importable pan_mod) and the panel-error polynomials; without
    it every ray sits on an error-free panel (the zero-offset case of ff_sim.py:61-62 with the stand-in layout)."""
    tg = ff_tele_geo(freq, offset)
    theta, phi = ff_fan(n_scan)
    plan = FarFieldPlan(tg, theta, phi, data_x, data_y, e_rec, device)
    pixels = ff_pixels(tg, res_arcmin, num)
    rank, world = dist_info()
    # ranks take contiguous blocks of the pixel list ordered so that mirror partners (+az, el), (-az, el) are
    # neighbours: the kernel traces such a pair once (a pixel's value does not depend on the order or the sharding)
    order = mirror_pair_order(pixels) if world > 1 else np.arange(len(pixels))
    lo, hi = shard_bounds(len(pixels), rank, world)
    mine = pixels[order[lo:hi]]
    if hi > lo and panel_models is not None:
        amp, phase, st = plan.run_panel_models(mine, panel_models[0], panel_models[1])
    elif hi > lo:
        amp, phase, st = plan.run(mine)
    else:
        amp, phase, st = np.zeros(0), np.zeros(0), 0
    _lib.raise_on_status(st, strict)
    plan.close()
    if world == 1:
        return pixels, amp, phase
    import torch

    dev = _gather_device(plan.ctx.device)
    local = torch.from_numpy(np.stack([amp, phase], axis=1)).to(dev)
    full = gather_blocks(local, len(pixels))
    if full is None:
        return None
    full = full.cpu().numpy()
    out = np.empty_like(full)
    out[order] = full
    return pixels, np.ascontiguousarray(out[:, 0]), np.ascontiguousarray(out[:, 1])


def save_map_txt(path, c0, c1, amp, phase):
    """The reference's output format (holo_sim.py:131-139, ff_sim.py:181-189): 4 columns, %.18e."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    np.savetxt(path, np.c_[np.ravel(c0), np.ravel(c1), np.ravel(amp), np.ravel(phase)])
