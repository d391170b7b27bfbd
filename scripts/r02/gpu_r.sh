#!/bin/bash
# round 2, call R: K4s2 with lazy own-row shuffles (windows of 2 doubles per row) at 2 / 3 blocks per SM vs the eager form
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_lattice_large.py -m gpu -q -k "streaming" ) > gpurun_out/r02_r_pytest_eager.log 2>&1; tail -2 gpurun_out/r02_r_pytest_eager.log
( ODEINTR_STREAM2_LAZY=1 timeout 600 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_lattice_large.py -m gpu -q -k "2d or 2D or laplace or tiled2d or stream" ) > gpurun_out/r02_r_pytest_lazy.log 2>&1; tail -2 gpurun_out/r02_r_pytest_lazy.log
for cfg in "eager 2" "lazy 2" "lazy 3"; do set -- $cfg
  for m in 8192 16384; do
    echo "== stream2d $1 min_blocks=$2 M=$m"
    if [ "$1" = lazy ]; then export ODEINTR_STREAM2_LAZY=1; else unset ODEINTR_STREAM2_LAZY; fi
    ODEINTR_STREAM2_MIN_BLOCKS=$2 QUICK=1 M=$m STEPS=40 timeout 200 python scripts/bench_tiled2d.py 2>&1 | grep "cell-steps" | tail -1
  done
done > gpurun_out/r02_r_sweep_stream2d.log 2>&1
cat gpurun_out/r02_r_sweep_stream2d.log
