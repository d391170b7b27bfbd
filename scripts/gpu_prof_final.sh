#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("roofline", {k: d["roofline"][k] for k in ("bound", "achieved", "peak", "frac", "traffic")})
print("cpu", d["cpu_baseline"])
print("secondary", json.dumps(d["secondary"])[:900])
PY
python scripts/prof_tiled.py > gpurun_out/prof_tiled_plain.log 2>&1 || { cat gpurun_out/prof_tiled_plain.log; exit 1; }
cat gpurun_out/prof_tiled_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_tiled_tma.csv python scripts/prof_tiled.py > gpurun_out/ncu_tiled_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:odeintr_lattice_rk4_tiled_tma -s 1 -c 1 \
    -o gpurun_out/prof_tiled_tma_bistable python scripts/prof_tiled.py > gpurun_out/ncu_tiled_full.log 2>&1
echo "ncu full (bistable) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:odeintr_lattice_rk4_tiled_tma -s 5 -c 1 \
    -o gpurun_out/prof_tiled_tma_chain python scripts/prof_tiled.py > gpurun_out/ncu_tiled_full2.log 2>&1
echo "ncu full (chain) rc=$?"
python scripts/prof_lattice_ensemble.py > gpurun_out/prof_ens_plain.log 2>&1 || { cat gpurun_out/prof_ens_plain.log; exit 1; }
cat gpurun_out/prof_ens_plain.log
ncu --set full --clock-control none --import-source on -k regex:odeintr_lattice_ensemble -c 2 \
    -o gpurun_out/prof_lattice_ensemble python scripts/prof_lattice_ensemble.py > gpurun_out/ncu_ens_full.log 2>&1
echo "ncu full (ensemble) rc=$?"
ls -la gpurun_out/*.ncu-rep
