#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q -k "2d or laplace or tiled" > gpurun_out/pytest_gpu_lattice.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_lattice.log
timeout 600 python scripts/bench_tiled2d.py 2>&1 | tee gpurun_out/bench_tiled2d.log
for cfg in "ODEINTR_TILE2_MIN_BLOCKS=4" "ODEINTR_TILE2_MIN_BLOCKS=2" "ODEINTR_TILE2_ROWS=64 ODEINTR_TILE2_MIN_BLOCKS=2" "ODEINTR_TILE2_ROWS=16 ODEINTR_TILE2_MIN_BLOCKS=5"; do
  echo "== $cfg" | tee -a gpurun_out/bench_tiled2d.log
  env $cfg QUICK=1 timeout 300 python scripts/bench_tiled2d.py 2>&1 | tail -3 | tee -a gpurun_out/bench_tiled2d.log
done
