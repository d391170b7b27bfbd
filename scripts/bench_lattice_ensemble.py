"""On a B200: ensembles of wider systems (K5, one warp / block per trajectory): M trajectories of a 1-D bistable lattice
with per-trajectory parameters, rk4, observer every 100 steps.  Prints cell-steps/s and the FP64-issue fraction
(54 DP instructions per cell and step with run-time parameters: 4 x 10 RHS + 14 stage algebra)."""
import ctypes as C, os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
peak = C.c_double(); ms = C.c_double()
assert lib.odeintr_fp64_probe(1, 0, C.byref(peak), C.byref(ms)) == 0
for cells, m in ((32, 1 << 19), (128, 1 << 17), (1024, 1 << 14), (4096, 1 << 12), (8192, 1 << 11)):
    s = odeintr_b200.compile_sys_openmp("bis", BISTABLE_1D, BISTABLE_PARS, False, method="rk4", sys_dim=cells).system
    rng = np.random.default_rng(1)
    x0 = rng.integers(0, 2, (cells, m)).astype(np.float64)
    s.set_params(D=rng.uniform(0.05, 0.3, m), a=1.0, b=rng.uniform(0.3, 0.7, m))
    steps = 1000
    for rep in range(3):
        s.set_state(x0)
        assert lib.odeintr_integrate_const(s._h, 0.0, steps * 0.01, 0.01, 100, 1) == 0, _abi.last_error()
        t = s.last_kernel_ms()
    rate = cells * m * s.last_steps() / t * 1e3
    print("lattice ensemble: %5d cells x %7d trajectories, %d steps: %.2f ms -> %.4g cell-steps/s = %.2f of the FP64 issue peak; %s"
          % (cells, m, s.last_steps(), t, rate, rate * 54 / peak.value, s.kernel_attributes("odeintr_lattice_ensemble")), flush=True)
    s.close()
