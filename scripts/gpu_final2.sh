#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log

python scripts/prof_2d_d5.py > gpurun_out/prof_2d_d5_plain.log 2>&1 || { cat gpurun_out/prof_2d_d5_plain.log; }
cat gpurun_out/prof_2d_d5_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_2d_d5.csv python scripts/prof_2d_d5.py > gpurun_out/ncu_2d_d5_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"odeintr_lattice_rk4_tiled2d|odeintr_lattice_dopri5_tiled" -s 1 -c 3 \
    -o gpurun_out/prof_2d_d5 python scripts/prof_2d_d5.py > gpurun_out/ncu_2d_d5_full.log 2>&1
echo "ncu full rc=$?"
