#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q -k "2d or laplace" > gpurun_out/pytest_gpu_lattice.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_lattice.log
QUICK=1 timeout 300 python scripts/bench_tiled2d.py 2>&1 | tail -3
ODEINTR_TILE2_ROWS=64 ODEINTR_TILE2_MIN_BLOCKS=2 QUICK=1 timeout 300 python scripts/bench_tiled2d.py 2>&1 | tail -3
ODEINTR_TILE2_ROWS=16 ODEINTR_TILE2_MIN_BLOCKS=4 QUICK=1 timeout 300 python scripts/bench_tiled2d.py 2>&1 | tail -3
