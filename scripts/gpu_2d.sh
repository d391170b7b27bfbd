#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_lattice_ensemble.py -m gpu -x -q > gpurun_out/pytest_gpu_lattice.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_lattice.log
timeout 600 python scripts/bench_tiled2d.py 2>&1 | tee gpurun_out/bench_tiled2d.log
