#!/bin/bash
# round 2, call Y3: K5a with the right-hand side inlined (the divisions fold now): bench secondary, then the tests
mkdir -p gpurun_out
python - > gpurun_out/r02_y3_k5a.log 2>&1 <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import bench
from odeintr_b200 import _abi
lib = _abi.load()
r = bench.lattice_ensemble_adaptive_secondary(lib, 1.857e13)
print(json.dumps({k: r[k] for k in ("value", "ms_per_pass", "parity")}))
PY
cat gpurun_out/r02_y3_k5a.log
( timeout 600 python -m pytest tests/test_gpu_lattice_ensemble.py -m gpu -q ) > gpurun_out/r02_y3_pytest.log 2>&1; tail -3 gpurun_out/r02_y3_pytest.log
