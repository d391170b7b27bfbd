#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep_tiled.log
run() { echo "== $*" | tee -a gpurun_out/sweep_tiled.log; env "$@" CHECK=0 timeout 200 python scripts/bench_tiled.py 2>&1 | grep -v "^$" | tail -n +1 | tee -a gpurun_out/sweep_tiled.log; }
CHECK=1 timeout 300 python scripts/bench_tiled.py 2>&1 | tee -a gpurun_out/sweep_tiled.log
run CASES=bistable ODEINTR_TILE_MIN_BLOCKS=5
run CASES=bistable ODEINTR_TILE_MIN_BLOCKS=3
run CASES=bistable ODEINTR_TILE_CPT=8 ODEINTR_TILE_MIN_BLOCKS=2
run CASES=bistable ODEINTR_TILE_CPT=2 ODEINTR_TILE_MIN_BLOCKS=6
run CASES=chain ODEINTR_TILE_MIN_BLOCKS=2
run CASES=chain ODEINTR_TILE_MIN_BLOCKS=4
run CASES=chain ODEINTR_TILE_CPT=2 ODEINTR_TILE_MIN_BLOCKS=4
run CASES=chain ODEINTR_TILE_CPT=2 ODEINTR_TILE_MIN_BLOCKS=5
