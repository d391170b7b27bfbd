#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
python bench.py --steps 5 > gpurun_out/bench_n1_quick.json 2> gpurun_out/bench_n1_quick.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1_quick.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
for k, v in (d.get("secondary") or {}).items():
    print("secondary", k, v.get("value"), (v.get("roofline") or {}).get("frac"), v.get("error"))
PY
