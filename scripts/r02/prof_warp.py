"""ncu target: a few rk4 steps of the 2^28-cell bistable lattice (K4w) and of the 2^28-state oscillator chain."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import BISTABLE_1D, BISTABLE_PARS, CHAIN, CHAIN_GLOBALS, CHAIN_PARS
lib = _abi.load()
n = 1 << 28
which = sys.argv[1:] or ["bistable", "chain"]
if "bistable" in which:
    s = odeintr_b200.compile_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n).system
    s.set_state((np.arange(n) % 7 < 3).astype(np.float64))
    for _ in range(2):
        assert lib.odeintr_integrate_const(s._h, 0.0, 0.08, 0.01, 1, 0) == 0, _abi.last_error()
    print("bistable", s.last_kernel_ms() / 8, "ms/step", flush=True)
    s.close()
if "chain" in which:
    s = odeintr_b200.compile_sys_openmp("c", CHAIN, CHAIN_PARS, True, method="rk4", sys_dim=n, globals=CHAIN_GLOBALS).system
    x = np.zeros(n); x[:n // 2] = np.sin(2 * np.pi * 8 * np.arange(n // 2) / (n // 2))
    s.set_state(x)
    for _ in range(2):
        assert lib.odeintr_integrate_const(s._h, 0.0, 0.08, 0.01, 1, 0) == 0, _abi.last_error()
    print("chain", s.last_kernel_ms() / 8, "ms/step", flush=True)
    s.close()
