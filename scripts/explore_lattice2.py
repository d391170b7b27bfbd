"""Scratch on a B200: 2-D periodic lattice through laplace2D (the reference's own large-system example shape) and
adaptive dopri5 on a 1-D lattice."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
M = 16384
n = M * M
s = odeintr_b200.compile_sys_openmp("b2d", BISTABLE_2D, BISTABLE_PARS, True, method="rk4", sys_dim=n).system
print("2d", s.kernel_attributes("odeintr_lattice_stage"), flush=True)
rng = np.random.default_rng(1)
s.set_state(rng.integers(0, 2, n).astype(np.float64))
for rep in range(3):
    assert lib.odeintr_integrate_const(s._h, 0.0, 0.5, 0.01, 1, 0) == 0, _abi.last_error()
    ms = s.last_kernel_ms(); st = s.last_steps()
    print("2-D bistable %dx%d rk4 steps=%d %.2f ms -> %.4g cell-steps/s, %.1f GB/s algorithmic" % (M, M, st, ms, n * st / ms * 1e3, 128.0 * n * st / ms / 1e6), flush=True)
s.close()
n = 1 << 26
s = odeintr_b200.compile_sys_openmp("b1d", BISTABLE_1D, BISTABLE_PARS, True, sys_dim=n).system  # default rk5_i
print("combo", s.kernel_attributes("odeintr_lattice_combo"), flush=True)
x0 = rng.uniform(0, 1, n)
for rep in range(2):
    s.set_state(x0)
    at = np.array([0.0, 1.0, 5.0, 10.0])
    assert lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) == 0, _abi.last_error()
    ms = s.last_kernel_ms(); stt = s.stats()
    tries = int(stt["accepted"][0] + stt["rejected"][0])
    print("1-D bistable 2^26 rk5_i bistable_at: %.2f ms, accepted %d rejected %d -> %.2f ms/try, %.4g cell-tries/s" % (ms, stt["accepted"][0], stt["rejected"][0], ms / tries, n * tries / ms * 1e3), flush=True)
