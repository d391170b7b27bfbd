#!/bin/bash
# round 2, call N: K5a tests (fixed routing, noinline RHS), K4ds2 error-norm loads early vs late
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_lattice_ensemble.py -m gpu -q -k "adaptive" ) > gpurun_out/r02_n_pytest.log 2>&1
tail -25 gpurun_out/r02_n_pytest.log
( M=8192 METHODS=rk5_a timeout 600 python scripts/r02/bench_dopri5_2d.py ) > gpurun_out/r02_n_dopri5_2d_early.log 2>&1
( ODEINTR_STREAM5_LATE_NORM=1 M=8192 METHODS=rk5_a timeout 600 python scripts/r02/bench_dopri5_2d.py ) > gpurun_out/r02_n_dopri5_2d_late.log 2>&1
cat gpurun_out/r02_n_dopri5_2d_early.log gpurun_out/r02_n_dopri5_2d_late.log
