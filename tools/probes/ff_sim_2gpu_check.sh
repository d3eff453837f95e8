#!/bin/bash
# 1-GPU vs 2-GPU scripts/ff_sim.py (and holo_sim.py) on small sizes: the sharded maps must equal the single-GPU ones.
set -e
cd "$(dirname "$0")/../.."
export PYTHONPATH=$PWD
T=$(mktemp -d)
python scripts/holo_sim.py 140 --res 9 --n-scan 30 --out $T/nf1.txt > /dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/holo_sim.py 140 --res 9 --n-scan 30 --out $T/nf2.txt > /dev/null 2>&1
python scripts/ff_sim.py 140 0.25 6 1.32 --nf-input $T/nf1.txt --n-scan 40 --out $T/ff1.txt > /dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 scripts/ff_sim.py 140 0.25 6 1.32 --nf-input $T/nf1.txt --n-scan 40 --out $T/ff2.txt > /dev/null 2>&1
python - <<PY
import numpy as np
for k in ("nf", "ff"):
    a = np.loadtxt("$T/%s1.txt" % k); b = np.loadtxt("$T/%s2.txt" % k)
    print(k, a.shape, "coords equal", np.array_equal(a[:, :2], b[:, :2]), "max |d amp|/amp %.2e" % np.max(np.abs(a[:, 2] - b[:, 2]) / a[:, 2]),
          "max |d phase| %.2e" % np.max(np.abs(a[:, 3] - b[:, 3])), "bit-identical rows %d / %d" % ((a == b).all(axis=1).sum(), len(a)))
PY
