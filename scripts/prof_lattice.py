"""Short lattice run for ncu: bistable 1-D chain, 2^28 cells, 5 rk4 steps (20 stage launches)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import BISTABLE_1D, BISTABLE_PARS
lib = _abi.load()
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
code = odeintr_b200.compile_sys_openmp("bistable", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n)
s = code.system
s.set_state((np.arange(n) % 7 < 3).astype(np.float64))
rc = lib.odeintr_integrate_const(s._h, 0.0, 0.05, 0.01, 1, 0)
assert rc == 0, _abi.last_error()
print("steps", s.last_steps(), "ms", s.last_kernel_ms())
