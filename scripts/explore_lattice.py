"""Scratch exploration on a B200: lattice rk4 (config 5) achieved HBM bandwidth."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
for name, src, pars, glob, logn in (("bistable", BISTABLE_1D, BISTABLE_PARS, "", 28), ("chain", CHAIN, CHAIN_PARS, CHAIN_GLOBALS, 28),
                                    ("bistable", BISTABLE_1D, BISTABLE_PARS, "", 24)):
    n = 1 << logn
    code = odeintr_b200.compile_sys_openmp(name, src, pars, True, method="rk4", sys_dim=n, globals=glob,
                                           halo=(1 if os.environ.get("TILED") else None))
    s = code.system
    print(name, n, s.kernel_attributes("odeintr_lattice_stage"), flush=True)
    if name == "chain":
        L = n // 2
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(L) / L), np.zeros(L)])
    else:
        x0 = (np.arange(n) % 7 < 3).astype(np.float64)
    s.set_state(x0)
    for rep in range(3):
        rc = lib.odeintr_integrate_const(s._h, 0.0, 1.0, 0.01, 1, 0)
        assert rc == 0, _abi.last_error()
        ms = s.last_kernel_ms()
        steps = s.last_steps()
        print("%s n=2^%d steps=%d %.2f ms -> %.4g cell-steps/s, %.1f GB/s at 128 B/cell/step" %
              (name, logn, steps, ms, n * steps / ms * 1e3, 128.0 * n * steps / ms * 1e3 / 1e9), flush=True)
    s.close()
