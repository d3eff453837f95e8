"""Whole-map plans and multi-GPU sharding.

Replaces the reference's MPI master/worker fan-out (scripts/holo_sim.py:87-189,
scripts/ff_sim.py:135-227, scripts/beamsim_set.sh): the scan-pixel list is cut into contiguous
blocks, one per rank (one process per GPU under torchrun), every rank runs ONE fused kernel over
its block, and a single gather (NCCL over NVLink; gloo in the CPU tests) brings (amp, phase) to
rank 0.  Pixels are independent, so there is no other exchange step.
"""
import ctypes
import functools
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
# Below this much work (scan pixels x launched rays) a map runs on rank 0 alone: the reference-sized maps take a
# fraction of a millisecond on one GPU, less than a collective costs.
SHARD_MIN_PIXEL_RAYS = 2.0e8


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


def _fold(col):
    """|col| rounded to 1e-9 of its extent: np.linspace grids are symmetric about 0 only to a few ulp."""
    scale = float(np.max(np.abs(col))) if len(col) else 0.0
    return np.round(np.abs(col) * (1e9 / scale if scale > 0 else 1.0))


def mirror_pair_order(pixels, fold_second=False):
    """Permutation of an (az, el) -- or near-field (x, z) -- pixel list that puts (+az, el) next to (-az, el).
    fold_second: the four near-field positions (+-x, +-z) next to each other."""
    pixels = np.asarray(pixels, dtype=np.float64)
    if fold_second:
        return np.lexsort((np.signbit(pixels[:, 0]), np.signbit(pixels[:, 1]), _fold(pixels[:, 1]), _fold(pixels[:, 0])))
    return np.lexsort((np.signbit(pixels[:, 0]), pixels[:, 1], _fold(pixels[:, 0])))


def shard_groups(pixels, world_size, fold_second=False):
    """Which pixels every rank computes: list of `world_size` index arrays into `pixels`.
    The mirror groups -- (+-az, el), or (+-x, +-z) with fold_second -- stay whole (the kernel traces a group once) and
    are dealt to the ranks cyclically in (|az|, el) order, so every rank gets an even sample of the map: the cost of
    a pixel varies smoothly over the map (how much of the fan reaches the E_rec grid), contiguous blocks do not
    balance.  Replaces the master/worker queue of scripts/ff_sim.py:135-227 by a static deal."""
    pixels = np.asarray(pixels, dtype=np.float64).reshape(-1, 2)
    order = mirror_pair_order(pixels, fold_second)
    if len(order) == 0:
        return [order.copy() for _ in range(world_size)]
    k0 = _fold(pixels[:, 0])[order]
    k1 = (_fold(pixels[:, 1]) if fold_second else pixels[:, 1])[order]
    first = np.ones(len(order), dtype=bool)
    first[1:] = (k0[1:] != k0[:-1]) | (k1[1:] != k1[:-1])
    gid = np.cumsum(first) - 1
    return [order[gid % world_size == r] for r in range(world_size)]


class _digest:
    """Cache key over the CONTENTS of arrays: hashes a cheap fingerprint (shapes + 64-bit word sums), compares the full
    contents (kept as private copies) on a hit -- exact, and ~10x faster than a cryptographic digest of the tables."""

    __slots__ = ("arrays", "_hash")

    def __init__(self, *arrays):
        self.arrays = [np.array(a, copy=True, order="C") for a in arrays]
        fp = []
        for a in self.arrays:
            raw = a.reshape(-1).view(np.uint8)
            n8 = raw.size // 8
            fp.append((a.shape, a.dtype.str, int(raw[: n8 * 8].view(np.uint64).sum(dtype=np.uint64)) if n8 else 0,
                       raw[n8 * 8:].tobytes()))
        self._hash = hash(tuple(fp))

    def __hash__(self):
        return self._hash

    @staticmethod
    def _same_bytes(a, b):
        """bit-for-bit equality of two C-contiguous arrays (NaN-safe), compared 8 bytes at a time"""
        ra, rb = a.reshape(-1).view(np.uint8), b.reshape(-1).view(np.uint8)
        n8 = ra.size // 8 * 8
        return (np.array_equal(ra[:n8].view(np.uint64), rb[:n8].view(np.uint64)) and ra[n8:].tobytes() == rb[n8:].tobytes())

    def __eq__(self, other):
        return (isinstance(other, _digest) and self._hash == other._hash and len(self.arrays) == len(other.arrays)
                and all(a.shape == b.shape and a.dtype == b.dtype and self._same_bytes(a, b)
                        for a, b in zip(self.arrays, other.arrays)))


class _PlanCache:
    """Plans (device tables, scratch buffers) are kept across calls: a map call with the same geometry, fan, grid and
    receiver field reuses them instead of uploading everything again.  LRU; an entry that falls out (or
    `clear_plan_cache()`) is freed when its last user lets go of it."""

    def __init__(self, size=8):
        self.size, self.items = size, {}

    def get(self, key, make):
        hit = self.items.pop(key, None)
        if hit is None:
            hit = make()
            while len(self.items) >= self.size:
                self.items.pop(next(iter(self.items)))
        self.items[key] = hit
        return hit

    def clear(self):
        self.items.clear()


_plans = _PlanCache()
_rx_fields = _PlanCache(16)
_sharded = _PlanCache(2)


def clear_plan_cache():
    _sharded.clear()
    _plans.clear()
    _rx_fields.clear()


class ShardedMap:
    """One map dealt over the ranks of the process group (one process per GPU).  Every rank keeps its pixels in HBM;
    `step()` enqueues, on the current CUDA stream, the rank's fused map kernel and -- with more than one rank -- the
    single collective of the path: the NCCL gather of (amp, phase, ray status) to rank 0.  `fetch()` brings the
    gathered map to rank 0's host memory and puts it back in the caller's pixel order.
    `plan`: NearFieldPlan / FarFieldPlan (channels = outputs per pixel)."""

    def __init__(self, plan, pixels, fold_second, rank=None, world=None):
        import torch

        if rank is None:
            rank, world = dist_info()
        self.plan, self.rank, self.world = plan, rank, world
        pixels = _lib.as_f64(np.reshape(pixels, (-1, 2)))
        self.npix = len(pixels)
        self.parts = shard_groups(pixels, world, fold_second)
        self.nch = plan.channels
        mine = self.parts[rank]
        self.n_local = len(mine)
        self.blk = max(len(p) for p in self.parts)
        self.dev = _gather_device(plan.ctx.device, default_cuda=True)   # CPU only under the gloo tests
        self.cuda = self.dev.type == "cuda"
        pin = (lambda t: t.pin_memory()) if self.cuda else (lambda t: t)
        self.flat = 2 * self.nch * self.blk + 1                       # [amp | phase] of this rank, then the status slot
        self.h_pix = pin(torch.from_numpy(pixels[mine].copy()))
        self.d_pix = self.h_pix.to(self.dev, non_blocking=True) if self.cuda else self.h_pix.clone()
        self.d_buf = torch.zeros(self.flat, dtype=torch.float64, device=self.dev)
        n = self.nch * self.n_local
        self.d_amp, self.d_phase, self.d_status = self.d_buf[:n], self.d_buf[n:2 * n], self.d_buf[self.flat - 1:]
        self.d_all = self.h_all = self.gather_list = None
        if rank == 0:
            self.d_all = torch.zeros((world, self.flat), dtype=torch.float64, device=self.dev) if world > 1 else None
            self.gather_list = list(self.d_all.unbind(0)) if world > 1 else None
            self.h_all = pin(torch.zeros((world, self.flat), dtype=torch.float64))
            # where every output element sits in the gathered buffer: the un-permute of fetch() is one gather per array
            src = np.empty((self.nch, self.npix), dtype=np.int64)
            for r, idx in enumerate(self.parts):
                if len(idx):
                    src[:, idx] = r * self.flat + np.arange(self.nch)[:, None] * len(idx) + np.arange(len(idx))[None, :]
            self._src_amp = src
            self._src_phase = src + self.nch * np.array([len(p) for p in self.parts], dtype=np.int64)[self._rank_of()][None, :]

    def _rank_of(self):
        """rank that computes each pixel"""
        owner = np.empty(self.npix, dtype=np.int64)
        for r, idx in enumerate(self.parts):
            owner[idx] = r
        return owner

    def _stream(self):
        import torch

        return torch.cuda.current_stream() if self.cuda else None

    def upload(self):
        """pinned host -> HBM copy of this rank's pixel list (part of an end-to-end step)"""
        self.d_pix.copy_(self.h_pix, non_blocking=True)

    def step(self):
        import torch

        if self.n_local:
            self.plan.run_device(self.d_pix, self.d_amp, self.d_phase, self._stream().cuda_stream if self.cuda else 0, self.d_status)
        if self.world > 1:
            import torch.distributed as dist

            dist.gather(self.d_buf, self.gather_list, dst=0)

    def fetch(self):
        """Rank 0: (amp [channels, npix], phase [channels, npix], status) in the caller's pixel order; None elsewhere."""
        import torch

        if self.rank != 0:
            if self.cuda:
                self._stream().synchronize()
            return None
        self.h_all.copy_(self.d_all if self.world > 1 else self.d_buf.unsqueeze(0), non_blocking=True)
        if self.cuda:
            self._stream().synchronize()
        h = self.h_all.numpy()
        flat = h.reshape(-1)
        amp, phase = flat[self._src_amp], flat[self._src_phase]
        status = 0
        for r, idx in enumerate(self.parts):
            if len(idx):
                status |= int(h[r, self.flat - 1:].view(np.int32)[0])
        return amp, phase, status

    def bytes_h2d(self):
        return int(self.h_pix.numel() * 8)

    def bytes_d2h(self):
        return int(self.world * self.flat * 8) if self.rank == 0 else 0


def _use_ranks(n_pixel_rays, world):
    """How many ranks a map of this size is dealt to (all of them, or rank 0 alone)."""
    return world if (world > 1 and n_pixel_rays >= SHARD_MIN_PIXEL_RAYS) else 1


def _finish_status(st, strict, sharded):
    """Strict mode raises on every rank together (the reference's workers raise where brentq does)."""
    if strict and sharded:
        import torch
        import torch.distributed as dist

        t = torch.tensor([int(st or 0)], dtype=torch.int64, device=_gather_device(int(os.environ.get("LOCAL_RANK", "0"))))
        dist.broadcast(t, src=0)
        st = int(t.item())
    _lib.raise_on_status(st or 0, strict)


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

    def run_device(self, d_src, d_amp, d_phase, stream=0, d_status=None):
        """torch CUDA float64 tensors ([npix,2], [channels*npix], [channels*npix]); asynchronous on `stream`.
        d_status: optional 8-byte device slot that receives the launch's ray status (else: the plan's sticky word)."""
        npix = d_src.shape[0]
        if d_status is None:
            _lib.check(_lib.load().solat_nf_map_device(self.handle, d_src.data_ptr(), npix, d_amp.data_ptr(),
                                                       d_phase.data_ptr(), stream))
        else:
            _lib.check(_lib.load().solat_nf_map_device_st(self.handle, d_src.data_ptr(), npix, d_amp.data_ptr(),
                                                          d_phase.data_ptr(), d_status.data_ptr(), stream))

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


@functools.lru_cache(maxsize=8)
def _holo_pixels(res, sweep, shift):
    rx = np.linspace(-sweep / 2 + shift, sweep / 2 + shift, res)
    xx, zz = np.meshgrid(rx, rx, indexing="ij")
    return np.stack([xx.ravel(), zz.ravel()], axis=1)


def holo_pixels(res=51, sweep=100, shift=0):
    """Source-position grid of scripts/holo_sim.py:90-108: row index i*res+j <-> (x_i, z_j)."""
    return _holo_pixels(int(res), float(sweep), float(shift)).copy()


def _with_digest(field):
    opl, w = field
    return opl, w, _digest(opl, w)


def _cached_receiver_field(tg, p_rx, n_scan_rx, n_scan):
    """(rx_opl, rx_w, digest of both) -- the digest keys the plan cache without hashing the tables on every call"""
    key = ("rx", _lib.context().device, tg.n_si, tg.n_vac, tg.th_fwhp, tg.y_ap, tuple(float(v) for v in p_rx), n_scan_rx, n_scan)
    return _rx_fields.get(key, lambda: _with_digest(receiver_field(tg, p_rx, n_scan_rx, n_scan)))


def near_field_plan(tg, k_list, rx_opl, rx_w, n_scan, device=None, field_digest=None):
    """NearFieldPlan from the plan cache (same geometry, wavenumbers and receiver field -> the same plan)."""
    dev = _lib.context(device).device
    key = ("nf", dev, tg.n_si, tg.n_vac, tg.th_fwhp, tg.y_ap, int(n_scan), tuple(float(k) for k in np.atleast_1d(k_list)),
           field_digest or _digest(rx_opl, rx_w))
    return _plans.get(key, lambda: NearFieldPlan(tg, k_list, rx_opl, rx_w, n_scan, device))


def _run_map(plan, pixels, fold_second, n_pixel_rays, strict):
    """Run `plan` over `pixels` on the ranks the map is worth dealing to.  Rank 0: (amp [channels, npix], phase);
    other ranks: None."""
    rank, world = dist_info()
    use = _use_ranks(n_pixel_rays, world)
    if use == 1:
        if rank != 0:
            return None
        amp, phase, st = plan.run(pixels)
        _lib.raise_on_status(st, strict)
        return np.reshape(amp, (plan.channels, -1)), np.reshape(phase, (plan.channels, -1))
    sm = _sharded.get((id(plan), _digest(pixels), world, fold_second), lambda: ShardedMap(plan, pixels, fold_second, rank, world))
    sm.upload()
    sm.step()
    res = sm.fetch()
    _finish_status(res[2] if res else 0, strict, True)
    return None if res is None else (res[0], res[1])


def holo_map(freqs, res=51, sweep=100, shift=0, n_scan=100, n_scan_rx=None, p_rx=(0, 0, 0), device=None,
             strict=False):
    """Near-field holography map(s) for one band.  `freqs`: scalar or list of GHz values that share
    the band geometry (holo_sim.py:40-45).  Under torch.distributed a map large enough to be worth it is dealt over
    the ranks (shard_groups), smaller ones run on rank 0.
    Returns on rank 0: (pixels [npix,2], amp [nfreq,npix], phase [nfreq,npix]); None elsewhere."""
    freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    geos = [holo_tele_geo(f, n_scan) for f in freqs]
    if len({g.y_ap for g in geos}) != 1:
        raise ValueError("frequencies of different bands (y_ap differs) need separate maps")
    tg = geos[0]
    rank, world = dist_info()
    pixels = holo_pixels(res, sweep, shift)
    work = len(pixels) * float(n_scan) ** 2
    if _use_ranks(work, world) == 1 and rank != 0:
        return None
    # the feed taper differs per band only; one geometry -> one rx field
    rx_opl, rx_w, dig = _cached_receiver_field(tg, p_rx, n_scan_rx, n_scan)
    plan = near_field_plan(tg, [g.k for g in geos], rx_opl, rx_w, n_scan, device, dig)
    out = _run_map(plan, pixels, True, work, strict)
    if out is None:
        return None
    return pixels, out[0], out[1]


def receiver_offset_sweep(freq, rx_positions, res=51, sweep=100, shift=0, n_scan=100, device=None, strict=False):
    """Near-field maps for many receiver feed positions at one frequency (the quant_systematics sweep,
    BASELINE config 4): the source rays are traced once per pixel and integrated against every receiver
    table.  Single process (shard the receiver list over ranks for more GPUs).
    Returns (pixels [npix,2], amp [n_rx,npix], phase [n_rx,npix])."""
    tg = holo_tele_geo(float(freq), n_scan)
    pos = _lib.as_f64(np.reshape(rx_positions, (-1, 3)))
    key = ("rxs", _lib.context(device).device, tg.n_si, tg.n_vac, tg.th_fwhp, tg.y_ap, int(n_scan), _digest(pos))
    rx_opl, rx_w, dig = _rx_fields.get(key, lambda: _with_digest(receiver_fields(tg, pos, n_scan, device)))
    plan = near_field_plan(tg, [tg.k], rx_opl, rx_w, n_scan, device, dig)
    pixels = holo_pixels(res, sweep, shift)
    amp, phase, st = plan.run(pixels)
    _lib.raise_on_status(st, strict)
    return pixels, amp[:, 0, :], phase[:, 0, :]


def _gather_device(device_index, default_cuda=False):
    """Device the collectives run on: the rank's GPU under NCCL, the CPU under gloo (CPU tests)."""
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return torch.device("cuda", device_index) if dist.get_backend() == "nccl" else torch.device("cpu")
    return torch.device("cuda", device_index) if default_cuda else torch.device("cpu")


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


@functools.lru_cache(maxsize=8)
def _ff_fan(n_scan):
    return np.linspace(np.pi - 0.00005, np.pi + 0.00005, n_scan), np.linspace(0, np.pi, n_scan)


def ff_fan(n_scan=100):
    """Polar launch fan of scripts/ff_sim.py:64-66."""
    th, ph = _ff_fan(int(n_scan))
    return th.copy(), ph.copy()


@functools.lru_cache(maxsize=8)
def _ff_pixels(az0, el0, res_arcmin, n):
    de_ang = res_arcmin / 60 * np.pi / 180
    azs = np.linspace(-n * de_ang + az0, n * de_ang + az0, 2 * n + 1)
    els = np.linspace(-n * de_ang + el0, n * de_ang + el0, 2 * n + 1)
    aa, ee = np.meshgrid(azs, els, indexing="ij")
    return np.stack([aa.ravel(), ee.ravel()], axis=1)


def ff_pixels(tele_geo, res_arcmin, num):
    """(az, el) grid of scripts/ff_sim.py:94,121-128: row index i*len+j <-> (az_i, el_j)."""
    return _ff_pixels(float(tele_geo.az0), float(tele_geo.el0), float(res_arcmin), int(num)).copy()


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
        self.channels = 1
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

    def run_device(self, d_azel, d_amp, d_phase, stream=0, d_status=None):
        """torch CUDA float64 tensors ([npix,2], [npix], [npix]); asynchronous on `stream`.
        d_status: optional 8-byte device slot that receives the launch's ray status (else: the plan's sticky word)."""
        if d_status is None:
            _lib.check(_lib.load().solat_ff_map_device(self.handle, d_azel.data_ptr(), d_azel.shape[0], d_amp.data_ptr(),
                                                       d_phase.data_ptr(), stream))
        else:
            _lib.check(_lib.load().solat_ff_map_device_st(self.handle, d_azel.data_ptr(), d_azel.shape[0], d_amp.data_ptr(),
                                                          d_phase.data_ptr(), d_status.data_ptr(), stream))

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

    def run_panel_layout(self, azel, layout1, layout2, panel_model1, panel_model2):
        """Fused map with panel models, the panels identified on the device: layout1 / layout2 are the
        solat_apert_field.PanelLayout of M1 / M2.  Per ray the reference's two passes (ff_sim.py:208-209)."""
        azel = _lib.as_f64(np.reshape(azel, (-1, 2)))
        npix = azel.shape[0]
        amp, phase = np.empty(npix), np.empty(npix)
        if npix == 0:
            return amp, phase, 0
        pm1, pm2 = _lib.as_f64(panel_model1), _lib.as_f64(panel_model2)
        if pm1.ndim != 2 or pm2.ndim != 2 or pm1.shape[0] != 10 or pm2.shape[0] != 10:
            raise ValueError("panel models must be (10, n_panels)")
        keep = []

        def c_layout(lay, pm):
            rect = _lib.as_f64(lay.rect)
            col = np.ascontiguousarray(lay.model_columns(pm), dtype=np.int32)
            keep.extend([rect, col])
            return _lib.PanelLayoutC(_lib.dptr(rect), col.ctypes.data_as(_lib.c_ip), len(col))

        l1, l2 = c_layout(layout1, pm1), c_layout(layout2, pm2)
        st = _lib.c_i(0)
        _lib.check(_lib.load().solat_ff_map_panel_layout_host(
            self.handle, _lib.dptr(azel), npix, ctypes.byref(l1), ctypes.byref(l2), _lib.dptr(pm1), pm1.shape[1], _lib.dptr(pm2),
            pm2.shape[1], _lib.dptr(amp), _lib.dptr(phase), ctypes.byref(st)))
        return amp, phase, st.value

    def run_panel_models(self, azel, panel_model1, panel_model2, backend=None, batch=64):
        """ff_sim.py:204-221 for a list of pixels with the panel models of pan_mod.  A backend that describes its panels
        as rectangles (`panel_layouts()`, see solat_apert_field.PanelLayout) is evaluated on the device
        (run_panel_layout).  Any other backend is opaque: mirror coordinates on the GPU -> its `identify_panel` on the
        host, pixel by pixel exactly as the reference calls it (saf:587-621) -> fused GPU map; `batch` pixels share one
        round trip."""
        from . import solat_apert_field as saf

        backend = backend or saf.panel_backend()
        if hasattr(backend, "panel_layouts"):
            lay1, lay2 = backend.panel_layouts()
            return self.run_panel_layout(azel, lay1, lay2, panel_model1, panel_model2)
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


def far_field_plan(tg, theta, phi, data_x, data_y, e_rec, device=None):
    """FarFieldPlan from the plan cache (same telescope, fan, E_rec grid and values -> the same plan)."""
    g = _lib.ff_geom(tg)
    key = ("ff", _lib.context(device).device, bytes(g), float(tg.k),
           _digest(theta, phi, np.real(data_x), np.real(data_y), np.asarray(e_rec, dtype=np.complex128)))
    return _plans.get(key, lambda: FarFieldPlan(tg, theta, phi, data_x, data_y, e_rec, device))


def ff_map(freq, res_arcmin, num, offset, data_x, data_y, e_rec, n_scan=100, device=None, strict=False, panel_models=None):
    """Far-field beam map `ff_sim.py <freq> <res> <num> <offset>`.  Returns on rank 0
    (pixels [npix,2] (az, el), amp [npix], phase [npix]); None on the other ranks.
    Under torch.distributed the mirror pairs (+-az, el) are dealt cyclically over the ranks (shard_groups), every
    rank runs the fused kernels over its pixels in HBM and one NCCL gather brings (amp, phase, status) to rank 0;
    maps too small to be worth a collective run on rank 0 alone.
    `panel_models` = (pan_mod1, pan_mod2) of ff_sim.py:69-70 switches on panel identification through the panel
    backend (solat_apert_field.set_panel_backend / an importable pan_mod) and the panel-error polynomials; without
    it every ray sits on an error-free panel (the zero-offset case of ff_sim.py:61-62 with the stand-in layout)."""
    tg = ff_tele_geo(freq, offset)
    pixels = ff_pixels(tg, res_arcmin, num)
    rank, world = dist_info()
    work = len(pixels) * float(n_scan) ** 2
    if _use_ranks(work, world) == 1 and rank != 0:
        return None
    theta, phi = ff_fan(n_scan)
    plan = far_field_plan(tg, theta, phi, data_x, data_y, e_rec, device)
    if panel_models is not None:
        return _ff_map_panels(plan, pixels, panel_models, work, strict)
    out = _run_map(plan, pixels, False, work, strict)
    if out is None:
        return None
    return pixels, out[0][0], out[1][0]


def _ff_map_panels(plan, pixels, panel_models, work, strict):
    """ff_map with panel models: contiguous blocks of the pair-ordered list per rank, host-side gather."""
    rank, world = dist_info()
    use = _use_ranks(work, world)
    order = mirror_pair_order(pixels) if use > 1 else np.arange(len(pixels))
    lo, hi = shard_bounds(len(pixels), rank, use)
    if hi > lo:
        amp, phase, st = plan.run_panel_models(pixels[order[lo:hi]], panel_models[0], panel_models[1])
    else:
        amp, phase, st = np.zeros(0), np.zeros(0), 0
    if use == 1:
        _lib.raise_on_status(st, strict)
        return pixels, amp, phase
    import torch

    dev = _gather_device(plan.ctx.device)
    local = torch.from_numpy(np.stack([amp, phase, np.full(len(amp), float(st))], axis=1)).to(dev)
    full = gather_blocks(local, len(pixels))
    st_all = 0 if full is None else int(np.bitwise_or.reduce(full[:, 2].cpu().numpy().astype(np.int64)))
    _finish_status(st_all, strict, True)
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
