#!/bin/bash
# round 2, call E: full GPU suite, K4dw sweep + launch list, ncu captures of the changed kernels, smoke, bench at N = 1
mkdir -p gpurun_out
( time ODEINTR_TEST_REPORT=1 timeout 1800 python -m pytest tests -m gpu -q -s --durations=8 ) > gpurun_out/r02_e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_e_pytest.log
grep -v "close()" gpurun_out/r02_e_pytest.log | tail -25
for cpt in 5 7 9; do for mb in 1 2; do
  echo "== dopri5 warp cpt=$cpt min_blocks=$mb"; ODEINTR_WARP5_CPT=$cpt ODEINTR_WARP5_MIN_BLOCKS=$mb timeout 200 python scripts/bench_dopri5_lattice.py 2>&1 | grep "rk5_i 2" | tail -1
done; done > gpurun_out/r02_e_sweep_dopri5.log 2>&1
cat gpurun_out/r02_e_sweep_dopri5.log
( timeout 600 python __graft_entry__.py smoke ) > gpurun_out/r02_e_smoke.log 2>&1; echo "smoke rc=$?"; tail -9 gpurun_out/r02_e_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r02_e_bench_n1.json 2> gpurun_out/r02_e_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_e_bench_n1.err
python scripts/r02/prof_all.py > gpurun_out/r02_e_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_e_launches.csv python scripts/r02/prof_all.py > gpurun_out/r02_e_ncu_launches.log 2>&1
cat gpurun_out/r02_e_prof_plain.log
ncu --set full --clock-control none --import-source on -k regex:"warp_h1|adaptive_ensemble|plain_ensemble" -c 14 -o gpurun_out/r02_prof_all python scripts/r02/prof_all.py > gpurun_out/r02_e_ncu_full.log 2>&1
tail -3 gpurun_out/r02_e_ncu_full.log
