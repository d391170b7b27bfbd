"""Short runs for ncu: 2-D tiled rk4 (8192^2, 3 steps) and tiled dopri5 (2^26 cells, a few attempts)."""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
m = 8192
s = odeintr_b200.compile_sys_openmp("b2d", BISTABLE_2D, BISTABLE_PARS, True, method="rk4", sys_dim=m * m).system
s.set_state((np.arange(m * m) % 7 < 3).astype(np.float64))
assert lib.odeintr_integrate_const(s._h, 0.0, 0.03, 0.01, 1, 0) == 0, _abi.last_error()
print("2d halo", s.get_halo(), "steps", s.last_steps(), "ms", s.last_kernel_ms(), flush=True); s.close()
n = 1 << 26
s = odeintr_b200.compile_sys_openmp("b1d", BISTABLE_1D, BISTABLE_PARS, True, method="rk5_a", sys_dim=n).system
s.set_state(np.random.default_rng(1).uniform(0, 1, n))
at = np.array([0.0, 1.0])
assert lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) == 0, _abi.last_error()
print("dopri5 ms", s.last_kernel_ms(), "launches", s.last_launches(), s.stats()["accepted"][0], flush=True); s.close()
