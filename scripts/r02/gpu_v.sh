#!/bin/bash
# round 2, call V: bench.py at N = 1 after the parity / CPU-sample additions to the secondaries; launch list of the bench command
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r02_v_bench_n1.json 2> gpurun_out/r02_v_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_v_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_v_bench_n1.json").read().strip().splitlines()[-1])
for k, v in (d.get("secondary") or {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "error", "parity", "bench_seconds")}, (v.get("cpu_baseline") or {}).get("value"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/r02_v_ncu_bench.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_v_ncu_bench.log | cut -c1-200; wc -l gpurun_out/r02_v_launches_bench.csv
