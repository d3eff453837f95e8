"""cuobjdump -sass instruction histogram of the timed kernels of libsolat_b200.so -> profiles/r02_sass_histogram.md (dev tool)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

LIB = os.path.join(ROOT, "solat_optics_b200", "libsolat_b200.so")
WANT = ["ff_map_kernelILi0ELi768ELi1", "ff_map_kernelILi0ELi256ELi3", "ff_map_kernelILi2ELi256ELi3", "ff_prep_kernelILi64", "ff_prep_kernelILi256",
        "nf_map_kernelILi1ELi1ELi2", "nf_phase_kernel", "nf_contract_kernel", "nf_contract_dmma_kernel", "ff_pair_kernel"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = re.search(r"arch = (sm_\w+)", txt)
out = ["# SASS instruction histogram of the timed kernels (static counts, `cuobjdump -sass solat_optics_b200/libsolat_b200.so`)", "",
       "csrc sha256 `%s`, cubin arch `%s`.  Hand-written FP64 SIMT path: DFMA/DMUL/DADD dominate; `MUFU` = reciprocal / rsqrt seeds;"
       % (bench.csrc_hash()[:16], arch.group(1) if arch else "?"),
       "`DMMA` only in the optional tensor-core contraction; `LDGSTS` = its cp.async staging; no `UTC*MMA` / `UTMA*` (the path has no",
       "dense low-precision contraction and its operands are generated from indices, north_star).", ""]
cur, hist = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next((w for w in WANT if w in m.group(1)), None)
        if cur and cur in hist:
            cur = None          # first instantiation only
        if cur:
            hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if cur and m:
        hist[cur][m.group(1)] += 1
ops = ["DFMA", "DMUL", "DADD", "DSETP", "DMMA", "FFMA", "FMUL", "FADD", "MUFU", "F2F", "IMAD", "ISETP", "LOP3", "SEL", "LDC", "LDCU", "LDS", "STS", "LDG", "STG",
       "LDL", "STL", "LDGSTS", "ATOMS", "ATOMG", "SHFL", "BAR", "BRA", "CALL"]
out.append("| kernel | total | " + " | ".join(ops) + " |")
out.append("|---|---|" + "---|" * len(ops))
for k in WANT:
    if k in hist:
        h = hist[k]
        out.append("| `%s` | %d | " % (k, sum(h.values())) + " | ".join(str(h.get(o, 0)) for o in ops) + " |")
open(os.path.join(ROOT, "profiles", "r02_sass_histogram.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
