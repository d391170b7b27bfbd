#!/bin/bash
# GPU tests, then the tiled rk4 kernel at 2^28 cells, then a one-rank slab run through the sharded entry points.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_tiled.py 2>&1 | tee gpurun_out/bench_tiled.log
timeout 300 python scripts/run_sharded_lattice.py 27 50 2>&1 | tee gpurun_out/sharded1_tiled.log
TILED=0 timeout 300 python scripts/run_sharded_lattice.py 27 50 2>&1 | tee gpurun_out/sharded1_staged.log
