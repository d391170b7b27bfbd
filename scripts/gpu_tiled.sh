#!/bin/bash
# GPU tests, then the tiled rk4 kernel at 2^28 cells with a sweep over cells per thread.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for cpt in 4 2 8; do
  echo "CPT=$cpt"
  CHECK=$([ $cpt = 4 ] && echo 1 || echo 0) ODEINTR_TILE_CPT=$cpt timeout 300 python scripts/bench_tiled.py 2>&1 | tee -a gpurun_out/bench_tiled.log
done
