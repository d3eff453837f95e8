"""Whole-map wall times of the five BASELINE.json configurations through the public API (dev/measurement tool, GPU).

    python tools/full_maps.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/full_maps.py

Every map is computed twice (first call pays context / plan creation warm-up); the second call is timed with a
barrier + device synchronize on both sides, wall clock, max over ranks, gather to rank 0 included.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from solat_optics_b200 import maps

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def fence():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        fence()
        t0 = time.perf_counter()
        out = fn()
        fence()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        best = dt if best is None else min(best, dt)
    return best, out


def rx_sweep():
    """config 4: 25 receiver offsets x near-field maps at 150 GHz; receivers sharded over ranks"""
    ax = np.linspace(-60, 60, 5)
    pos = [(x, 0.0, z) for x in ax for z in ax]
    lo, hi = maps.shard_bounds(len(pos), rank, world)
    if hi > lo:
        # single-process call per rank: drop the process group view so that the pixel list is not sharded again
        return maps.receiver_offset_sweep(150.0, pos[lo:hi], device=local)
    return None


results = []


def report(name, pixel_rays, dt):
    if rank == 0:
        r = {"config": name, "n_gpus": world, "wall_s": dt, "pixel_rays": pixel_rays, "pixel_rays_per_s": pixel_rays / dt}
        results.append(r)
        print(json.dumps(r), flush=True)


# config 1: default near-field map, 140 GHz
dt, out = timed(lambda: maps.holo_map(140.0, device=local))
report("1: near-field 140 GHz, 51x51 px x 100^2 rays", 2601 * 10 ** 4, dt)
nf = out
# config 3: frequency sweep (two bands: 90 GHz alone, 150/220/280 share one trace)
dt, _ = timed(lambda: (maps.holo_map(90.0, device=local), maps.holo_map([150.0, 220.0, 280.0], device=local)))
report("3: near-field sweep 90 | 150,220,280 GHz", 2 * 2601 * 10 ** 4, dt)
# config 4
dt, _ = timed(rx_sweep)
report("4: 25 receiver offsets x near-field map (150 GHz)", 25 * 2601 * 10 ** 4, dt)
# config 2: far-field shipped-file geometry, E_rec = config-1 map
if world > 1:
    obj = [nf]
    dist.broadcast_object_list(obj, src=0)
    nf = obj[0]
px, amp, ph = nf
x, y, e = maps.near_field_to_e_rec(np.c_[px[:, 0], px[:, 1], amp[0], ph[0]])
dt, _ = timed(lambda: maps.ff_map(140, 0.25, 20, 1.32, x, y, e, device=local))
report("2: far-field 140 GHz 1.32 m, 41x41 px x 100^2 rays", 1681 * 10 ** 4, dt)
# config 5: 401 x 401 px x 10^6 rays, closed-form E_rec on a 101 x 101 grid
ax = np.linspace(-5, 5, 101)
yy, xx = np.meshgrid(ax, ax, indexing="ij")
e5 = np.exp(-(xx ** 2 + yy ** 2) / (2 * 1.5 ** 2)) * np.exp(1j * 0.3 * (xx ** 2 + yy ** 2))
n5 = int(os.environ.get("SOLAT_CFG5_NUM", "200"))
dt, out5 = timed(lambda: maps.ff_map(140, 0.05, n5, 1.32, xx, yy, e5, n_scan=1000, device=local), reps=1 if world == 1 else 2)
report("5: far-field %dx%d px x 1000^2 rays" % (2 * n5 + 1, 2 * n5 + 1), (2 * n5 + 1) ** 2 * 10 ** 6, dt)
if rank == 0:
    a5 = out5[1]
    print("config 5 peak amp %.6g at pixel %d, min %.3g" % (a5.max(), int(a5.argmax()), a5.min()))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/full_maps_n%d.json" % world, "w") as f:
        json.dump(results, f, indent=1)
if world > 1:
    dist.destroy_process_group()
