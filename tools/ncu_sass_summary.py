"""Summarise an `ncu --page source --csv` dump: instruction mix by opcode and stall reasons (dev tool)."""
import csv
import sys
from collections import Counter

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
inst = Counter(); samples = Counter(); stalls = Counter()
tot_i = 0; tot_s = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr) or r[ix["Instructions Executed"]] == "Instructions Executed":
        continue
    if not r[ix["Address"]].startswith("0x"):
        continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    op = parts[0]
    if op.startswith("@"):
        op = parts[1]
    op = op.split(".")[0] if not op.startswith(("DFMA", "DMUL", "DADD", "MUFU")) else op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    inst[op] += n; samples[op] += s; tot_i += n; tot_s += s
    for c in stall_cols:
        stalls[c] += int(r[ix[c]] or 0)
print("total warp-instructions", tot_i, "samples", tot_s)
print("%-10s %12s %7s %8s" % ("opcode", "inst", "inst%", "sample%"))
for op, n in inst.most_common(28):
    print("%-10s %12d %6.1f%% %7.1f%%" % (op, n, 100.0 * n / tot_i, 100.0 * samples[op] / max(tot_s, 1)))
fp64 = sum(n for op, n in inst.items() if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("FP64-pipe instructions: %d (%.1f%%)" % (fp64, 100.0 * fp64 / tot_i))
print("stall reasons (all samples):")
ts = sum(stalls.values())
for c, n in stalls.most_common(10):
    print("  %-28s %6.1f%%" % (c, 100.0 * n / max(ts, 1)))
