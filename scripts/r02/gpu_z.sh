#!/bin/bash
# round 2, call Z: ncu of K5a (adaptive wide ensemble)
mkdir -p gpurun_out
python scripts/r02/prof_k5a.py > gpurun_out/r02_z_k5a_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"ensemble_dopri5_dense" -s 1 -c 1 -o gpurun_out/r02_prof_k5a python scripts/r02/prof_k5a.py > gpurun_out/r02_z_ncu.log 2>&1
cat gpurun_out/r02_z_k5a_plain.log; tail -2 gpurun_out/r02_z_ncu.log
