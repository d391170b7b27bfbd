#!/bin/bash
# Round-end style validation on one B200: smoke, GPU tests, both bench arms, secondary configs.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; wc -l gpurun_out/bench_n1.json; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("roofline", {k: d["roofline"][k] for k in ("bound", "achieved", "peak", "frac", "traffic")})
print("cpu", d["cpu_baseline"])
for k, v in (d.get("secondary") or {}).items():
    print("secondary", k, v.get("value"), (v.get("roofline") or {}).get("frac"), v.get("error"))
PY
timeout 900 python scripts/bench_configs.py 3 5 > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err; echo "configs rc=$?"; cut -c1-400 gpurun_out/bench_configs.jsonl
