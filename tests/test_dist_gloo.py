"""world_size-2 test of the multi-rank path on CPU (gloo): contiguous pixel blocks, one gather to rank 0,
ragged last block.  The per-rank 'kernel' here is a deterministic function of the pixel coordinates; the
CUDA maps are covered by the -m gpu tests (sharded == unsharded bit for bit)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_map(pixels):
    return np.stack([np.hypot(pixels[:, 0], pixels[:, 1]) + 1.0, np.arctan2(pixels[:, 1], pixels[:, 0] + 1e-3)], axis=1)


def _worker(rank, world, port, n_side, q, ordered=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    from solat_optics_b200 import maps

    assert maps.dist_info() == (rank, world)
    assert maps._gather_device(0) == torch.device("cpu")
    pixels = maps.holo_pixels(n_side, 100, 0)
    lo, hi = maps.shard_bounds(len(pixels), rank, world)
    # ordered: what holo_map / ff_map do -- blocks of the list sorted mirror partners next to each other, un-permuted on rank 0
    order = maps.mirror_pair_order(pixels, fold_second=True) if ordered else np.arange(len(pixels))
    local = torch.from_numpy(_fake_map(pixels[order[lo:hi]]))
    full = maps.gather_blocks(local, len(pixels))
    if rank == 0:
        out = np.empty_like(full.numpy())
        out[order] = full.numpy()
        q.put(out)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_side,ordered", [(2, 5, False), (2, 4, False), (3, 5, False), (2, 7, True), (3, 6, True)])
def test_sharded_gather_matches_single_rank(world, n_side, ordered):
    sys.path.insert(0, ROOT)
    from solat_optics_b200 import maps

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_side, q, ordered)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    np.testing.assert_array_equal(got, _fake_map(maps.holo_pixels(n_side, 100, 0)))


class _FakePlan:
    """Stands in for a NearFieldPlan / FarFieldPlan on a box without a GPU: same run_device signature, a deterministic
    function of the pixel coordinates, `channels` outputs per pixel, a status word in the 8-byte slot."""

    class ctx:
        device = 0

    def __init__(self, channels):
        self.channels = channels

    def run_device(self, d_pix, d_amp, d_phase, stream=0, d_status=None):
        px = d_pix.numpy()
        f = _fake_map(px)
        n = len(px)
        for c in range(self.channels):
            d_amp[c * n:(c + 1) * n] = torch.from_numpy(f[:, 0] + c)
            d_phase[c * n:(c + 1) * n] = torch.from_numpy(f[:, 1] - c)
        d_status.numpy().view(np.int32)[:] = (3 if (px[:, 0] == px[:, 0].max()).any() and px[:, 0].max() == 50.0 else 0, 0)


def _sharded_worker(rank, world, port, n_side, channels, fold, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    from solat_optics_b200 import maps

    pixels = maps.holo_pixels(n_side, 100, 0)
    sm = maps.ShardedMap(_FakePlan(channels), pixels, fold)
    assert not sm.cuda and sm.n_local == len(maps.shard_groups(pixels, world, fold)[rank])
    sm.upload()
    sm.step()
    res = sm.fetch()
    if rank == 0:
        q.put(res)
    else:
        assert res is None
    # strict mode: every rank learns the status rank 0 gathered
    try:
        maps._finish_status(res[2] if res else 0, True, True)
        raised = False
    except ValueError:
        raised = True
    assert raised
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_side,channels,fold", [(2, 5, 1, False), (3, 6, 2, True), (2, 7, 3, True), (3, 4, 1, False)])
def test_sharded_map_deals_groups_and_unpermutes(world, n_side, channels, fold):
    """ShardedMap (the multi-GPU path of holo_map / ff_map) on CPU ranks: cyclic deal of whole mirror groups, one
    gather with ragged per-rank counts, status OR, un-permute on rank 0."""
    sys.path.insert(0, ROOT)
    from solat_optics_b200 import maps

    pixels = maps.holo_pixels(n_side, 100, 0)
    parts = maps.shard_groups(pixels, world, fold)
    assert sorted(np.concatenate(parts).tolist()) == list(range(len(pixels)))
    key = lambda p: (round(abs(p[0]), 6), round(abs(p[1]), 6) if fold else p[1])
    owner = {}
    for r, idx in enumerate(parts):               # a mirror group never straddles two ranks
        for p in pixels[idx]:
            assert owner.setdefault(key(p), r) == r
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 4 * (2 if fold else 1)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, n_side, channels, fold, q)) for r in range(world)]
    for p in procs:
        p.start()
    amp, phase, status = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _fake_map(pixels)
    for c in range(channels):
        np.testing.assert_array_equal(amp[c], ref[:, 0] + c)
        np.testing.assert_array_equal(phase[c], ref[:, 1] - c)
    assert status == 3
