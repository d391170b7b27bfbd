#!/bin/bash
# round 2, call B: whole GPU suite on the new kernels (K4w warp kernel, overlapped shard_run, K6 order, exact division,
# sharded adaptive dopri5), then the K4w / K4t comparison, the K6 measurement and the K2 / K3 register sweep
mkdir -p gpurun_out
( time ODEINTR_TEST_REPORT=1 timeout 1800 python -m pytest tests -m gpu -q -s --durations=12 ) > gpurun_out/r02_b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_b_pytest.log
grep -v "^close()\|^\.close()" gpurun_out/r02_b_pytest.log | tail -60
( time timeout 600 python scripts/r02/bench_warp.py 50 ) > gpurun_out/r02_b_bench_warp.log 2>&1
cat gpurun_out/r02_b_bench_warp.log
( time timeout 600 python scripts/r02/bench_k6.py ) > gpurun_out/r02_b_bench_k6.log 2>&1
cat gpurun_out/r02_b_bench_k6.log
( time timeout 900 python scripts/r02/bench_k23.py ) > gpurun_out/r02_b_bench_k23.log 2>&1
cat gpurun_out/r02_b_bench_k23.log
