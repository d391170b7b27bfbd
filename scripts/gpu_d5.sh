#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_golden.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/pytest_gpu_lattice.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_lattice.log
timeout 400 python scripts/bench_dopri5_lattice.py 2>&1 | tee gpurun_out/bench_dopri5_lattice.log
ODEINTR_NO_TILED=1 timeout 400 python scripts/bench_dopri5_lattice.py 2>&1 | awk 'NR%3!=1' | tee -a gpurun_out/bench_dopri5_lattice.log
for cfg in "ODEINTR_TILE5_CPT=4 ODEINTR_TILE5_MIN_BLOCKS=2" "ODEINTR_TILE5_CPT=3 ODEINTR_TILE5_MIN_BLOCKS=3" "ODEINTR_TILE5_CPT=6 ODEINTR_TILE5_MIN_BLOCKS=2"; do
  echo "== $cfg" | tee -a gpurun_out/bench_dopri5_lattice.log
  env $cfg timeout 200 python scripts/bench_dopri5_lattice.py 2>&1 | grep -v sha1 | awk 'NR%2==0' | tee -a gpurun_out/bench_dopri5_lattice.log
done
