"""Short adaptive-ensemble runs for ncu: VdP rk5_i (config 3) and Robertson rosenbrock4 (config 4), 2^20 instances."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import VDP, ROBERTSON, ROBERTSON_PARS
lib = _abi.load()
m = 1 << 20
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("vdp", "both"):
    code = odeintr_b200.compile_sys("vpol", VDP, "mu", method="rk5_i", atol=1e-8, rtol=1e-8)
    s = code.system
    lo = (C.c_double * 1)(0.5); hi = (C.c_double * 1)(2.0)
    assert lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi) == 0
    times = np.arange(0, 20.0001, 0.2)
    s.set_state(np.repeat(np.array([2.0, 0.0])[:, None], m, axis=1))
    assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 0.01) == 0, _abi.last_error()
    st = s.stats()
    print("vdp ms %.3f accepted %d rejected %d" % (s.last_kernel_ms(), st["accepted"].sum(), st["rejected"].sum()))
if which in ("rob", "both"):
    code = odeintr_b200.compile_implicit("rob", ROBERTSON, ROBERTSON_PARS)
    s = code.system
    lo = (C.c_double * 3)(0.02, 1e4, 3e7); hi = (C.c_double * 3)(0.08, 1e4, 3e7)
    assert lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi) == 0
    times = 10 ** np.linspace(-5, 5, 41)
    s.set_state(np.repeat(np.array([1.0, 0.0, 0.0])[:, None], m, axis=1))
    assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 1e-6) == 0, _abi.last_error()
    st = s.stats()
    print("rob ms %.3f accepted %d rejected %d" % (s.last_kernel_ms(), st["accepted"].sum(), st["rejected"].sum()))
