#!/bin/bash
# round 2, call U (final): full GPU suite, smoke, bench at N = 1
mkdir -p gpurun_out
( time ODEINTR_TEST_REPORT=1 timeout 2400 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/r02_u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_u_pytest.log
grep -v "close()" gpurun_out/r02_u_pytest.log | tail -28
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/r02_u_smoke.log 2>&1; echo "smoke rc=$?"; tail -16 gpurun_out/r02_u_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r02_u_bench_n1.json 2> gpurun_out/r02_u_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_u_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_u_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"])
for k, v in (d.get("secondary") or {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "error", "parity", "bench_seconds")}, (v.get("roofline") or {}).get("frac"))
PY
