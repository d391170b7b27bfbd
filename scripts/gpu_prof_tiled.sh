#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_tiled.py > gpurun_out/prof_tiled_plain.log 2>&1 || { cat gpurun_out/prof_tiled_plain.log; exit 1; }
cat gpurun_out/prof_tiled_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_tiled.csv python scripts/prof_tiled.py > gpurun_out/ncu_tiled_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:odeintr_lattice_rk4_tiled_interior -s 1 -c 1 \
    -o gpurun_out/prof_tiled_bistable python scripts/prof_tiled.py > gpurun_out/ncu_tiled_full.log 2>&1
echo "ncu full (bistable) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:odeintr_lattice_rk4_tiled_interior -s 5 -c 1 \
    -o gpurun_out/prof_tiled_chain python scripts/prof_tiled.py > gpurun_out/ncu_tiled_full2.log 2>&1
echo "ncu full (chain) rc=$?"
ls -la gpurun_out/*.ncu-rep
