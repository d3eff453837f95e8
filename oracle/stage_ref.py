"""Stage the UNMODIFIED Python reference for timing on the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  `/root/reference` does not exist on the GPU box, so the reference's
CPU path could not be timed there.  This recipe copies the seven modules of the reference package that the hot
path uses, byte for byte, from `/root/reference/solat_optics/` into `oracle/_ref/solat_optics/` -- a directory that is
git-ignored (no reference source enters the history) but travels with the gpurun snapshot like the built `.so` files.
`bench.py --impl reference` and `oracle/ref_bench.py` import it from there through `oracle/ref_loader.py` (same four
shims as the golden-vector generator).  Run by `__graft_entry__.build()` whenever `/root/reference` is present.

    python oracle/stage_ref.py
"""
import hashlib
import os
import shutil

SRC = os.environ.get("SOLAT_REFERENCE_SRC", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(_HERE, "_ref")
MODULES = ["far_field_latrt.py", "latrt_geo.py", "ot_geo.py", "phase_correct.py", "ray_trace.py", "ray_trace_int.py",
           "solat_apert_field.py"]   # ray_trace.py: imported at module level by ray_trace_int.py:16


def stage(verbose=False):
    src = os.path.join(SRC, "solat_optics")
    if not os.path.isdir(src):
        return False
    dst = os.path.join(DST, "solat_optics")
    os.makedirs(dst, exist_ok=True)
    lines = []
    for m in MODULES:
        shutil.copyfile(os.path.join(src, m), os.path.join(dst, m))
        lines.append("%s  %s" % (hashlib.sha256(open(os.path.join(dst, m), "rb").read()).hexdigest(), m))
    with open(os.path.join(DST, "SHA256SUMS"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print("staged %d reference modules into %s" % (len(MODULES), dst))
    return True


if __name__ == "__main__":
    if not stage(verbose=True):
        raise SystemExit("reference not found at %s" % SRC)
