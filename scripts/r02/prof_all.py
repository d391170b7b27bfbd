"""ncu target of round 2: a few launches of every kernel that changed -- K4w (bistable, chain; two steps per pass),
K4dw (dopri5 attempt on a 2^26-cell lattice), K2 / K3 (10^6-instance sweeps), K1 (Lorenz rk4, 10^7 instances)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
which = sys.argv[1:] or ["warp", "dopri5", "adaptive", "rk4"]
if "warp" in which:
    n = 1 << 28
    s = odeintr_b200.compile_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n).system
    s.set_state((np.arange(n) % 7 < 3).astype(np.float64))
    assert lib.odeintr_integrate_const(s._h, 0.0, 0.08, 0.01, 1, 0) == 0, _abi.last_error()
    print("bistable", s.last_kernel_ms() / 8, "ms/step", s.last_launches(), flush=True); s.close()
    s = odeintr_b200.compile_sys_openmp("c", CHAIN, CHAIN_PARS, True, method="rk4", sys_dim=n, globals=CHAIN_GLOBALS).system
    x = np.zeros(n); x[:n // 2] = np.sin(2 * np.pi * 8 * np.arange(n // 2) / (n // 2))
    s.set_state(x)
    assert lib.odeintr_integrate_const(s._h, 0.0, 0.08, 0.01, 1, 0) == 0, _abi.last_error()
    print("chain", s.last_kernel_ms() / 8, "ms/step", s.last_launches(), flush=True); s.close()
if "dopri5" in which:
    n = 1 << 26
    s = odeintr_b200.compile_sys_openmp("b5", BISTABLE_1D, BISTABLE_PARS, True, sys_dim=n).system
    s.set_state(np.random.default_rng(1).uniform(0, 1, n))
    at = np.array([0.0, 1.0, 5.0])
    assert lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) == 0, _abi.last_error()
    print("dopri5 lattice", s.last_kernel_ms(), "ms", s.last_launches(), flush=True); s.close()
if "adaptive" in which:
    m = 1 << 20
    s = odeintr_b200.compile_sys("vpol", VDP, "mu", method="rk5_i", atol=1e-8, rtol=1e-8).system
    lo = (C.c_double * 1)(0.5); hi = (C.c_double * 1)(2.0)
    lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi)
    times = np.arange(0, 20.0001, 0.2)
    s.set_state(np.repeat(np.array([2.0, 0.0])[:, None], m, axis=1))
    assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 0.01) == 0
    st = s.stats(); print("vdp", s.last_kernel_ms(), "ms tries", int(st["accepted"].sum() + st["rejected"].sum()), flush=True); s.close()
    s = odeintr_b200.compile_implicit("rob", ROBERTSON, ROBERTSON_PARS).system
    lo = (C.c_double * 3)(0.02, 1e4, 3e7); hi = (C.c_double * 3)(0.08, 1e4, 3e7)
    lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi)
    times = 10 ** np.linspace(-5, 5, 41)
    s.set_state(np.repeat(np.array([1.0, 0.0, 0.0])[:, None], m, axis=1))
    assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 1e-6) == 0
    st = s.stats(); print("robertson", s.last_kernel_ms(), "ms tries", int(st["accepted"].sum() + st["rejected"].sum()), flush=True); s.close()
if "rk4" in which:
    m = 10_000_000
    s = odeintr_b200.compile_sys("lorenz", LORENZ, LORENZ_PARS, True, method="rk4").system
    lo = (C.c_double * 3)(-20, -25, 5); hi = (C.c_double * 3)(20, 25, 45)
    lib.odeintr_fill_state_uniform(s._h, m, 20260101, 0, lo, hi)
    assert lib.odeintr_integrate_const(s._h, 0.0, 1.0, 0.001, 100, 1) == 0
    print("lorenz rk4", s.last_kernel_ms(), "ms", flush=True); s.close()
