"""Time the unmodified Python reference on the host cores (bench.py --impl reference).

MEASUREMENT INFRASTRUCTURE ONLY.  One worker process per core (the reference fans pixels out over MPI ranks the same
way, scripts/ff_sim.py:135-227 / beamsim_set.sh:7; mpi4py is not in this image).  Every worker imports the reference
from oracle/_ref (staged by oracle/stage_ref.py) through oracle/ref_loader.py and runs the reference's own worker body
(ff_sim.py:204-221 / holo_sim.py:157-185) for the pixels it is handed.
"""
import multiprocessing as mp
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_state = {}


def _init():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    if not os.path.isdir("/root/reference/solat_optics"):
        os.environ["SOLAT_REFERENCE_ROOT"] = os.path.join(_HERE, "_ref")
    import importlib

    import numpy as np

    rl = importlib.import_module("oracle.ref_loader")
    importlib.reload(rl)
    _state["rl"] = rl
    _state["ref"] = rl.load()
    ax = np.linspace(-5, 5, 101)
    yy, xx = np.meshgrid(ax, ax, indexing="ij")
    _state["e5"] = (xx, yy, np.exp(-(xx ** 2 + yy ** 2) / (2 * 1.5 ** 2)) * np.exp(1j * 0.3 * (xx ** 2 + yy ** 2)))


def _ff_pixel(job):
    freq, offset, n_scan, az, el = job
    rl, ref = _state["rl"], _state["ref"]
    x, y, e = _state["e5"]
    amp, phi = rl.ff_pixel(ref, freq, offset, n_scan, az, el, x, y, e)
    return float(amp), float(phi)


def _nf_pixel(job):
    freq, x, z, n_scan = job
    rl, ref = _state["rl"], _state["ref"]
    amp, phi = rl.nf_pixel(ref, freq, x, z, n_scan)
    return float(amp), float(phi)


class Pool:
    """ProcessPoolExecutor (spawn): a worker that dies in its initialiser breaks the pool loudly instead of hanging."""

    def __init__(self, workers):
        from concurrent.futures import ProcessPoolExecutor

        self.workers = workers
        self.pool = ProcessPoolExecutor(workers, mp_context=mp.get_context("spawn"), initializer=_init)

    def ff_pixels(self, jobs):
        return list(self.pool.map(_ff_pixel, jobs, chunksize=1))

    def nf_pixels(self, jobs):
        return list(self.pool.map(_nf_pixel, jobs, chunksize=1))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.pool.shutdown(wait=True, cancel_futures=True)
