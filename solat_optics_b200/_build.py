"""Build libsolat_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "solat_kernels.cu")
HDRS = [os.path.join(_HERE, "csrc", "solat_math.cuh"), os.path.join(os.path.dirname(_HERE), "include", "solat_b200.h")]
LIB = os.environ.get("SOLAT_B200_LIB") or os.path.join(_HERE, "libsolat_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsolat_b200.so cannot be built")


def stale():
    if os.environ.get("SOLAT_B200_LIB"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in [SRC] + HDRS)


def build(force=False, verbose=False, extra=()):
    """Compile the CUDA library if it is missing or older than its sources.  Safe under torchrun on a fresh checkout:
    ranks serialise on a file lock, the first one compiles into a temporary file and renames it into place, the others
    find the library fresh when they get the lock."""
    if os.environ.get("SOLAT_B200_LIB"):
        return LIB          # an explicitly chosen prebuilt variant is never rebuilt
    if not force and not stale():
        return LIB
    import fcntl

    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():
                return LIB
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ["-o", tmp, SRC]
            if verbose:
                print(" ".join(cmd))
            try:
                subprocess.check_call(cmd)
                os.replace(tmp, LIB)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB
