"""Scratch exploration on a B200: adaptive ensembles (configs 3 and 4) throughput."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import VDP, ROBERTSON, ROBERTSON_PARS, STIFF
lib = _abi.load()

def run_vdp(method, m, x0):
    code = odeintr_b200.compile_sys("vpol", VDP, "mu", method=method, atol=1e-8, rtol=1e-8)
    s = code.system
    print(method, s.kernel_attributes("odeintr_adaptive_ensemble"), flush=True)
    lo = (C.c_double * 1)(0.5); hi = (C.c_double * 1)(2.0)
    assert lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi) == 0, _abi.last_error()
    times = np.arange(0, 20.0001, 0.2)
    X = np.repeat(np.array(x0)[:, None], m, axis=1)
    for rep in range(2):
        s.set_state(X)
        t0 = time.time()
        rc = lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 0.01)
        assert rc == 0, _abi.last_error()
        ms = s.last_kernel_ms()
    st = s.stats()
    acc, rej = st["accepted"].sum(), st["rejected"].sum()
    print("vdp %s x0=%s m=%d kernel %.2f ms accepted %.4g (mean %.1f, min %d max %d) rejected %.4g -> %.4g accepted-steps/s, %.4g tries/s" %
          (method, x0, m, ms, acc, acc / m, st["accepted"].min(), st["accepted"].max(), rej, acc / ms * 1e3, (acc + rej) / ms * 1e3), flush=True)
    s.close()

def run_rob(m):
    code = odeintr_b200.compile_implicit("rob", ROBERTSON, ROBERTSON_PARS)
    s = code.system
    print("rosenbrock4", s.kernel_attributes("odeintr_adaptive_ensemble"), flush=True)
    lo = (C.c_double * 3)(0.02, 1e4, 3e7); hi = (C.c_double * 3)(0.08, 1e4, 3e7)
    assert lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi) == 0
    times = 10 ** np.linspace(-5, 5, 41)
    X = np.repeat(np.array([1.0, 0.0, 0.0])[:, None], m, axis=1)
    for rep in range(2):
        s.set_state(X)
        rc = lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 1e-6)
        assert rc == 0, _abi.last_error()
        ms = s.last_kernel_ms()
    st = s.stats()
    acc, rej = st["accepted"].sum(), st["rejected"].sum()
    print("robertson m=%d kernel %.2f ms accepted %.4g (mean %.1f min %d max %d) rejected %.4g -> %.4g accepted-steps/s" %
          (m, ms, acc, acc / m, st["accepted"].min(), st["accepted"].max(), rej, acc / ms * 1e3), flush=True)
    s.close()

for m in (1 << 20, 10_000_000):
    run_vdp("rk5_i", m, (2.0, 0.0))
run_vdp("rk5_i", 10_000_000, (1e-4, 1e-4))
run_vdp("rk5_a", 10_000_000, (2.0, 0.0))
for m in (1 << 20, 10_000_000):
    run_rob(m)
