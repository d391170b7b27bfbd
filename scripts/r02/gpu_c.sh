#!/bin/bash
# round 2, call C: bisect the computed_index failure, failed tests again, K4w sweep, ncu capture of K4w
mkdir -p gpurun_out
python scripts/r02/debug_computed_index.py > gpurun_out/r02_c_debug.log 2>&1; cat gpurun_out/r02_c_debug.log
( timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_order.py tests/test_gpu_implicit.py -m gpu -q -k "general_snippets or order or implicit or computed_index or robertson" ) > gpurun_out/r02_c_pytest.log 2>&1
tail -15 gpurun_out/r02_c_pytest.log
bash scripts/r02/sweep_warp.sh > gpurun_out/r02_c_sweep_warp.log 2>&1; cat gpurun_out/r02_c_sweep_warp.log
python scripts/r02/prof_warp.py > gpurun_out/r02_c_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:odeintr_lattice_rk4_warp -c 8 -o gpurun_out/r02_prof_warp python scripts/r02/prof_warp.py > gpurun_out/r02_c_ncu.log 2>&1
tail -5 gpurun_out/r02_c_ncu.log
