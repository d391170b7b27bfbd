#!/bin/bash
# round 2, call L: bsd on large systems (fixed), K4ds2 (streaming dopri5 attempt for M x M lattices): tests, timing vs staged
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_lattice.py -m gpu -q -k "bulirsch or general_snippets or streaming_dopri5 or tiny or adaptive_dopri5" ) > gpurun_out/r02_l_pytest.log 2>&1
tail -15 gpurun_out/r02_l_pytest.log
( M=8192 timeout 600 python scripts/r02/bench_dopri5_2d.py ) > gpurun_out/r02_l_dopri5_2d_stream.log 2>&1
( ODEINTR_NO_STREAM2D=1 M=8192 METHODS=rk5_i timeout 600 python scripts/r02/bench_dopri5_2d.py ) > gpurun_out/r02_l_dopri5_2d_staged.log 2>&1
cat gpurun_out/r02_l_dopri5_2d_stream.log gpurun_out/r02_l_dopri5_2d_staged.log
for rows in 128 192 256 384; do echo "rows=$rows"; ODEINTR_STREAM5_ROWS=$rows M=8192 METHODS=rk5_i timeout 300 python scripts/r02/bench_dopri5_2d.py | grep attempt | tail -1; done > gpurun_out/r02_l_dopri5_2d_rows.log 2>&1
cat gpurun_out/r02_l_dopri5_2d_rows.log
