#!/bin/bash
# round 2, call B: lattice tests on the new warp kernel + overlapped shard_run, K6 order tests, then the K4w / K4t
# comparison and the K6 measurement
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_lattice_large.py tests/test_gpu_lattice_ensemble.py tests/test_gpu_order.py -m gpu -q --durations=8 ) > gpurun_out/r02_b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_b_pytest.log
tail -25 gpurun_out/r02_b_pytest.log
( time timeout 600 python scripts/r02/bench_warp.py 50 ) > gpurun_out/r02_b_bench_warp.log 2>&1
cat gpurun_out/r02_b_bench_warp.log
( time timeout 600 python scripts/r02/bench_k6.py ) > gpurun_out/r02_b_bench_k6.log 2>&1
cat gpurun_out/r02_b_bench_k6.log
