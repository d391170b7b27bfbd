#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_golden.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/pytest_gpu_lattice.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_lattice.log
timeout 400 python scripts/bench_dopri5_lattice.py 2>&1 | tee gpurun_out/bench_dopri5_lattice.log
ODEINTR_NO_TILED=1 timeout 400 python scripts/bench_dopri5_lattice.py 2>&1 | tee -a gpurun_out/bench_dopri5_lattice.log
