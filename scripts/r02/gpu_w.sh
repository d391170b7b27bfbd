#!/bin/bash
# round 2, call W: the streaming kernels with several tasks per warp (capped grid)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_lattice.py -m gpu -q -k "several_tasks" ) > gpurun_out/r02_w_pytest.log 2>&1; tail -5 gpurun_out/r02_w_pytest.log
