#!/bin/bash
# round 2, call Y: K5a with the right-hand side called by value (stage vectors in registers): tests + bench secondary
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_lattice_ensemble.py -m gpu -q ) > gpurun_out/r02_y_pytest.log 2>&1; tail -4 gpurun_out/r02_y_pytest.log
python - > gpurun_out/r02_y_k5a.log 2>&1 <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import bench
from odeintr_b200 import _abi
lib = _abi.load()
r = bench.lattice_ensemble_adaptive_secondary(lib, 1.857e13)
print(json.dumps(r))
PY
cut -c1-700 gpurun_out/r02_y_k5a.log
