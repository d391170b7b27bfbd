"""Scratch: how much does per-instance step-count heterogeneity cost the thread-per-trajectory adaptive kernel?
Warp efficiency = sum of tries / (32 * sum over warps of the max tries in the warp)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
m = 1 << 20
def report(label, s, ms):
    st = s.stats(); tries = st["accepted"] + st["rejected"]
    w = tries.reshape(-1, 32)
    eff = tries.sum() / (32.0 * w.max(axis=1).sum())
    print("%s: %.2f ms, tries mean %.1f min %d max %d, warp efficiency %.3f, %.4g tries/s" % (label, ms, tries.mean(), tries.min(), tries.max(), eff, tries.sum() / ms * 1e3), flush=True)
s = odeintr_b200.compile_sys("lor", LORENZ, LORENZ_PARS, True, method="rk5_i", atol=1e-8, rtol=1e-8).system
lo = (C.c_double * 3)(-20, -25, 5); hi = (C.c_double * 3)(20, 25, 45)
times = np.arange(0, 10.0001, 0.1)
for rep in range(2):
    lib.odeintr_fill_state_uniform(s._h, m, 1, 0, lo, hi)
    assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 0.01) == 0
report("lorenz random ICs rk5_i T=10", s, s.last_kernel_ms())
s.close()
s = odeintr_b200.compile_sys("vpol", VDP, "mu", method="rk5_i", atol=1e-8, rtol=1e-8).system
rng = np.random.default_rng(0)
mus = rng.uniform(0.1, 10.0, m)  # shuffled, wide stiffness range: heterogeneous neighbours
s.set_params(mu=mus)
times = np.arange(0, 20.0001, 0.2)
X = np.repeat(np.array([2.0, 0.0])[:, None], m, axis=1)
for rep in range(2):
    s.set_state(X)
    assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 0.01) == 0
report("vdp shuffled mu in [0.1, 10]", s, s.last_kernel_ms())
s.set_params(mu=np.sort(mus))
for rep in range(2):
    s.set_state(X)
    assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 0.01) == 0
report("vdp sorted mu in [0.1, 10]", s, s.last_kernel_ms())
