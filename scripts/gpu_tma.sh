#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_golden.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/pytest_gpu_lattice.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_lattice.log
: > gpurun_out/bench_tma.log
echo "== TMA persistent" | tee -a gpurun_out/bench_tma.log
CHECK=1 timeout 300 python scripts/bench_tiled.py 2>&1 | tee -a gpurun_out/bench_tma.log
echo "== plain loads (ODEINTR_NO_TMA=1)" | tee -a gpurun_out/bench_tma.log
ODEINTR_NO_TMA=1 CHECK=0 timeout 300 python scripts/bench_tiled.py 2>&1 | tee -a gpurun_out/bench_tma.log
echo "== TMA, MIN_BLOCKS sweeps" | tee -a gpurun_out/bench_tma.log
for mb in 3 5; do echo "bistable MIN_BLOCKS=$mb"; CASES=bistable ODEINTR_TILE_MIN_BLOCKS=$mb CHECK=0 timeout 200 python scripts/bench_tiled.py 2>&1 | tail -2; done | tee -a gpurun_out/bench_tma.log
for mb in 2 4; do echo "chain MIN_BLOCKS=$mb"; CASES=chain ODEINTR_TILE_MIN_BLOCKS=$mb CHECK=0 timeout 200 python scripts/bench_tiled.py 2>&1 | tail -2; done | tee -a gpurun_out/bench_tma.log
timeout 300 python scripts/run_sharded_lattice.py 27 50 2>&1 | tee gpurun_out/sharded1_tiled.log
