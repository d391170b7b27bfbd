#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lattice_ensemble.py -m gpu -x -q > gpurun_out/pytest_gpu_ens.log 2>&1; echo "pytest ens rc=$?"; tail -15 gpurun_out/pytest_gpu_ens.log
timeout 300 python scripts/bench_lattice_ensemble.py 2>&1 | tee gpurun_out/bench_lattice_ensemble.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 gpurun_out/pytest_gpu.log
