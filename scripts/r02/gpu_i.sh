#!/bin/bash
# round 2, call I: K4s2 with rotating register rows (no window shifts): tests, timing, ncu
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_lattice_large.py -m gpu -q -k "2d or 2D or laplace or tiled2d or stream" ) > gpurun_out/r02_i_pytest.log 2>&1
tail -3 gpurun_out/r02_i_pytest.log
for m in 8192 16384; do for ahead in 0; do
  echo "== stream2d M=$m l2_ahead=$ahead"; ODEINTR_STREAM2_L2_AHEAD=$ahead QUICK=1 M=$m STEPS=40 timeout 200 python scripts/bench_tiled2d.py 2>&1 | grep "cell-steps" | tail -1
done; done > gpurun_out/r02_i_sweep_stream2d.log 2>&1
cat gpurun_out/r02_i_sweep_stream2d.log
cat > /tmp/prof2d.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
m = 8192
s = odeintr_b200.compile_sys_openmp("b2d", BISTABLE_2D, BISTABLE_PARS, True, method="rk4", sys_dim=m * m).system
s.set_state(np.random.default_rng(1).integers(0, 2, m * m).astype(np.float64))
assert lib.odeintr_integrate_const(s._h, 0.0, 0.04, 0.01, 1, 0) == 0
print("2d", s.last_kernel_ms() / 4, "ms/step", flush=True)
PY
python /tmp/prof2d.py > gpurun_out/r02_i_prof2d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stream2d" -s 1 -c 2 -o gpurun_out/r02_prof_stream2d_rot python /tmp/prof2d.py > gpurun_out/r02_i_ncu1.log 2>&1
cat gpurun_out/r02_i_prof2d_plain.log; tail -2 gpurun_out/r02_i_ncu1.log
