"""Short run for ncu: K5 ensemble of 1024-cell bistable lattices (2^13 trajectories, 200 steps) and of 32-cell ones."""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
for cells, m in ((1024, 1 << 13), (32, 1 << 18)):
    s = odeintr_b200.compile_sys_openmp("bis", BISTABLE_1D, BISTABLE_PARS, False, method="rk4", sys_dim=cells).system
    rng = np.random.default_rng(1)
    s.set_params(D=rng.uniform(0.05, 0.3, m), a=1.0, b=rng.uniform(0.3, 0.7, m))
    s.set_state(rng.integers(0, 2, (cells, m)).astype(np.float64))
    assert lib.odeintr_integrate_const(s._h, 0.0, 2.0, 0.01, 100, 1) == 0, _abi.last_error()
    print(cells, m, "steps", s.last_steps(), "ms", s.last_kernel_ms(), flush=True)
    s.close()
