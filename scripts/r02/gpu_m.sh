#!/bin/bash
# round 2, call M: adaptive ensembles of wider systems (K5a), 2-D row-block slabs, K4ds2 rk5 case; ncu of K4ds2
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_lattice_ensemble.py tests/test_gpu_lattice.py -m gpu -q -k "ensemble or emulated_ranks or streaming_dopri5" ) > gpurun_out/r02_m_pytest.log 2>&1
tail -25 gpurun_out/r02_m_pytest.log
cat > /tmp/prof2d5.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
m = 8192
s = odeintr_b200.compile_sys_openmp("b2d", BISTABLE_2D, BISTABLE_PARS, True, method="rk5_a", sys_dim=m * m).system
s.set_state(np.random.default_rng(1).uniform(0, 1, m * m))
assert lib.odeintr_integrate_adaptive(s._h, 0.0, 0.5, 0.1, 0) == 0, _abi.last_error()
st = s.stats()
print("2d dopri5", s.last_kernel_ms() / (st["accepted"][0] + st["rejected"][0]), "ms/attempt", flush=True)
PY
python /tmp/prof2d5.py > gpurun_out/r02_m_prof2d5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"dopri5_stream2d" -s 1 -c 1 -o gpurun_out/r02_prof_dopri5_stream2d python /tmp/prof2d5.py > gpurun_out/r02_m_ncu1.log 2>&1
cat gpurun_out/r02_m_prof2d5_plain.log; tail -2 gpurun_out/r02_m_ncu1.log
