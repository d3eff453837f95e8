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
