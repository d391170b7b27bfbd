#!/bin/bash
# round 2, call J: bs / bsd on large systems (tests), K4s2 with the unconditional prefetch load
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_lattice.py -m gpu -q -x -k "bulirsch or general_snippets or cash_karp" ) > gpurun_out/r02_j_pytest.log 2>&1
tail -30 gpurun_out/r02_j_pytest.log
( timeout 600 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_lattice_large.py -m gpu -q -k "2d or 2D or laplace or tiled2d or stream" ) > gpurun_out/r02_j_pytest2d.log 2>&1
tail -3 gpurun_out/r02_j_pytest2d.log
for m in 8192 16384; do
  echo "== stream2d M=$m"; QUICK=1 M=$m STEPS=40 timeout 200 python scripts/bench_tiled2d.py 2>&1 | grep "cell-steps" | tail -1
done > gpurun_out/r02_j_sweep_stream2d.log 2>&1
cat gpurun_out/r02_j_sweep_stream2d.log
