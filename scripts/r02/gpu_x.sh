#!/bin/bash
# round 2, call X: bench.py at N = 1 with the adaptive wide-ensemble secondary
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r02_x_bench_n1.json 2> gpurun_out/r02_x_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_x_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_x_bench_n1.json").read().strip().splitlines()[-1])
for k, v in (d.get("secondary") or {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "error", "parity", "bench_seconds")}, (v.get("roofline") or {}).get("frac"))
PY
