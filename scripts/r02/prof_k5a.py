"""Short K5a run for ncu: 592 trajectories x 1024 cells, default rk5_i, NAME_at at 4 times."""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
cells, m = 1024, 592
s = odeintr_b200.compile_sys_openmp("bisa", BISTABLE_1D, BISTABLE_PARS, False, sys_dim=cells).system
rng = np.random.default_rng(2)
x0 = rng.uniform(0, 1, (cells, m))
s.set_params(D=rng.uniform(0.05, 0.3, m), a=1.0, b=rng.uniform(0.3, 0.7, m))
at = np.array([0.0, 1.0, 5.0, 10.0])
for rep in range(2):
    s.set_state(x0)
    assert lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) == 0, _abi.last_error()
    st = s.stats()
    print("k5a", s.last_kernel_ms(), "ms", int(st["accepted"].sum() + st["rejected"].sum()), "attempts", flush=True)
