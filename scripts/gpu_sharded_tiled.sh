#!/bin/bash
# slab-sharded lattice on N GPUs (N = $1), tiled kernels, halo exchange every 8 steps
N=${1:-8}
LOG2=${2:-25}
mkdir -p gpurun_out
SPE=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/run_sharded_lattice.py $LOG2 200 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee gpurun_out/sharded${N}_tiled_tma_spe8.log
