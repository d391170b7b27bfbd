#!/bin/bash
# round 2, call Q: K4ds2 threads per block (8 / 10 / 12 warps per SM: 255 / 204 / 168 registers), K3 at 4 blocks
mkdir -p gpurun_out
for t in 256 320 384; do echo "threads=$t"; ODEINTR_STREAM5_THREADS=$t M=8192 METHODS=rk5_i timeout 300 python scripts/r02/bench_dopri5_2d.py | grep -v "^stream.*attempt" | tail -1; ODEINTR_STREAM5_THREADS=$t M=8192 METHODS=rk5_i timeout 300 python scripts/r02/bench_dopri5_2d.py | grep attempt | tail -1; done > gpurun_out/r02_q_dopri5_2d_threads.log 2>&1
cat gpurun_out/r02_q_dopri5_2d_threads.log
( timeout 600 python -m pytest tests/test_gpu_implicit.py tests/test_gpu_lattice.py -m gpu -q -k "implicit or robertson or stiff or streaming_dopri5" ) > gpurun_out/r02_q_pytest.log 2>&1; tail -3 gpurun_out/r02_q_pytest.log
