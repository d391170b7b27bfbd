#!/bin/bash
mkdir -p gpurun_out
for cfg in "ODEINTR_TILE5_CPT=2 ODEINTR_TILE5_MIN_BLOCKS=3" "ODEINTR_TILE5_CPT=4 ODEINTR_TILE5_MIN_BLOCKS=2" "ODEINTR_TILE5_CPT=4 ODEINTR_TILE5_MIN_BLOCKS=3" "ODEINTR_TILE5_CPT=3 ODEINTR_TILE5_MIN_BLOCKS=3" "ODEINTR_TILE5_CPT=1 ODEINTR_TILE5_MIN_BLOCKS=5" "ODEINTR_TILE5_CPT=2 ODEINTR_TILE5_MIN_BLOCKS=4"; do
  echo "== $cfg"
  env $cfg timeout 200 python scripts/bench_dopri5_lattice.py 2>&1 | grep -v sha1 | awk 'NR%2==0'
done | tee gpurun_out/sweep_dopri5.log
