#!/bin/bash
# round 2, call A: whole GPU suite with deviation report, smoke(), default bench at N = 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_a_gpu.txt 2>&1
free -g >> gpurun_out/r02_a_gpu.txt; nproc >> gpurun_out/r02_a_gpu.txt
( time ODEINTR_TEST_REPORT=1 timeout 1500 python -m pytest tests -m gpu -q -s --durations=15 ) > gpurun_out/r02_a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_a_pytest.log
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/r02_a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02_a_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r02_a_bench_n1.json 2> gpurun_out/r02_a_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02_a_bench_n1.err
tail -3 gpurun_out/r02_a_pytest.log; tail -3 gpurun_out/r02_a_smoke.log; tail -c 600 gpurun_out/r02_a_bench_n1.json
