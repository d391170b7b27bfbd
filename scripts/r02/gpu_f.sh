#!/bin/bash
# round 2, call F: the streaming 2-D kernel (K4s2): tests, full-size comparison with K4t2 / staged, segment-length sweep
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_lattice_large.py -m gpu -q -k "2d or 2D or laplace or tiled2d or stream" ) > gpurun_out/r02_f_pytest.log 2>&1
tail -8 gpurun_out/r02_f_pytest.log
( M=16384 STEPS=20 timeout 600 python scripts/bench_tiled2d.py ) > gpurun_out/r02_f_bench2d_stream.log 2>&1; cat gpurun_out/r02_f_bench2d_stream.log
( ODEINTR_NO_STREAM2D=1 QUICK=1 M=16384 STEPS=20 timeout 600 python scripts/bench_tiled2d.py ) > gpurun_out/r02_f_bench2d_block.log 2>&1; cat gpurun_out/r02_f_bench2d_block.log
for rows in 128 256 512 1024; do for mb in 1 2; do
  echo "== stream2d rows=$rows min_blocks=$mb"; ODEINTR_STREAM2_ROWS=$rows ODEINTR_STREAM2_MIN_BLOCKS=$mb QUICK=1 M=8192 STEPS=40 timeout 200 python scripts/bench_tiled2d.py 2>&1 | grep "cell-steps" | tail -1
done; done > gpurun_out/r02_f_sweep_stream2d.log 2>&1
cat gpurun_out/r02_f_sweep_stream2d.log
