#!/bin/bash
# slab-sharded lattice on N GPUs (N = $1): tiled kernels with an exchange every 8 / 1 steps, then the staged kernels
N=${1:-2}
LOG2=${2:-27}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/run_sharded_lattice.py $LOG2 100 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" ; }
SPE=8 run | tee gpurun_out/sharded${N}_tiled_spe8.log
SPE=1 run | tee gpurun_out/sharded${N}_tiled_spe1.log
TILED=0 run | tee gpurun_out/sharded${N}_staged.log
