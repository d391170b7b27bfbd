"""Short run for ncu: tiled rk4 on the 2^28-cell lattices (bistable: 1 field, chain: 2 fields), 4 steps each."""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
n = 1 << 28
for name, src, pars, glob in (("bistable", BISTABLE_1D, BISTABLE_PARS, ""), ("chain", CHAIN, CHAIN_PARS, CHAIN_GLOBALS)):
    s = odeintr_b200.compile_sys_openmp(name, src, pars, True, method="rk4", sys_dim=n, globals=glob).system
    s.set_state((np.arange(n) % 7 < 3).astype(np.float64))
    assert lib.odeintr_integrate_const(s._h, 0.0, 0.04, 0.01, 1, 0) == 0, _abi.last_error()
    print(name, "halo", s.get_halo(), "steps", s.last_steps(), "ms", s.last_kernel_ms(), "launches", s.last_launches(), flush=True)
    s.close()
