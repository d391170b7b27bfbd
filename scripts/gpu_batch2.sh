#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/explore_lattice.py > gpurun_out/explore_lattice.log 2>&1; echo rc=$?; tail -12 gpurun_out/explore_lattice.log
python scripts/prof_lattice.py 28 > gpurun_out/prof_lattice_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_lattice.csv python scripts/prof_lattice.py 28 > gpurun_out/ncu_lat_launches.log 2>&1
echo "ncu lattice launches rc=$?"; cat gpurun_out/prof_lattice_plain.log
python scripts/prof_adaptive.py both > gpurun_out/prof_adaptive_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:odeintr_adaptive_ensemble -c 2 -o gpurun_out/prof_adaptive_v1 python scripts/prof_adaptive.py both > gpurun_out/ncu_adaptive.log 2>&1
echo "ncu adaptive rc=$?"; cat gpurun_out/prof_adaptive_plain.log
