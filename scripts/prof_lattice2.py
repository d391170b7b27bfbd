"""Short runs for ncu: 1-D lattice rk4 (2^28), 2-D lattice rk4 (16384^2), 1-D lattice dopri5 (2^26)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
n = 1 << 28
s = odeintr_b200.compile_sys_openmp("bistable", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n).system
s.set_state((np.arange(n) % 7 < 3).astype(np.float64))
assert lib.odeintr_integrate_const(s._h, 0.0, 0.03, 0.01, 1, 0) == 0
print("1d steps", s.last_steps(), "ms", s.last_kernel_ms()); s.close()
M = 16384
s = odeintr_b200.compile_sys_openmp("b2d", BISTABLE_2D, BISTABLE_PARS, True, method="rk4", sys_dim=M * M).system
s.set_state((np.arange(M * M) % 7 < 3).astype(np.float64))
assert lib.odeintr_integrate_const(s._h, 0.0, 0.03, 0.01, 1, 0) == 0
print("2d steps", s.last_steps(), "ms", s.last_kernel_ms()); s.close()
n = 1 << 26
s = odeintr_b200.compile_sys_openmp("b1d", BISTABLE_1D, BISTABLE_PARS, True, sys_dim=n).system
s.set_state(np.random.default_rng(1).uniform(0, 1, n))
at = np.array([0.0, 0.5])
assert lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) == 0
print("dopri5 lattice ms", s.last_kernel_ms(), s.stats())
