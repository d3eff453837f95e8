"""Shared launcher glue for the CLI scripts: one process per GPU under torchrun, or a single process."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


def init_distributed():
    """Returns (rank, world).  Under torchrun (WORLD_SIZE > 1) initialises torch.distributed with NCCL,
    binding rank -> cuda:LOCAL_RANK; replaces `mpiexec -n N` of the reference (beamsim_set.sh:7)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend="nccl")
    return rank, world


def finish_distributed():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
