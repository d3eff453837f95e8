#!/usr/bin/env bash
# Frequency sweep driver (drop-in for the reference's scripts/beamsim_set.sh, which loops
# `mpiexec -n 70 python3 holo_sim.py $f`).  One process per GPU; frequencies of one band
# (<=120 GHz | >120 GHz) share a single ray trace when passed together.
GPUS=${GPUS:-$(nvidia-smi -L 2>/dev/null | wc -l)}
declare -a FREQ
FREQ=('150')
cd "$(dirname "$0")"
for f in "${FREQ[@]}"; do
    if [ "${GPUS:-1}" -gt 1 ]; then
        python3 -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 --master-port 29517 holo_sim.py "$f"
    else
        python3 holo_sim.py "$f"
    fi
done
