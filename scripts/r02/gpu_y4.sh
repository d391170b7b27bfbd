#!/bin/bash
# round 2, call Y4: K5a at two resident blocks per SM (128-register cap)
mkdir -p gpurun_out
ODEINTR_ENS_ADAPTIVE_MIN_BLOCKS=2 python - > gpurun_out/r02_y4_k5a.log 2>&1 <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import bench
from odeintr_b200 import _abi
lib = _abi.load()
r = bench.lattice_ensemble_adaptive_secondary(lib, 1.857e13)
print(json.dumps({k: r[k] for k in ("value", "ms_per_pass", "parity")}))
PY
cat gpurun_out/r02_y4_k5a.log
