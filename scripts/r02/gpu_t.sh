#!/bin/bash
# round 2, call T: K2 with the dopri5 tableau in constant memory vs literals; adaptive tests
mkdir -p gpurun_out
( timeout 600 python scripts/r02/bench_k2_constants.py ) > gpurun_out/r02_t_k2_constants.log 2>&1; cut -c1-260 gpurun_out/r02_t_k2_constants.log
( timeout 900 python -m pytest tests/test_gpu_adaptive.py tests/test_gpu_order.py -m gpu -q ) > gpurun_out/r02_t_pytest.log 2>&1; tail -3 gpurun_out/r02_t_pytest.log
