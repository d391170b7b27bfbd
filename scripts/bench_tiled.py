"""On a B200: the temporally blocked rk4 step (compile_sys_openmp(..., halo=1)) on BASELINE.json configs[4]-sized
lattices.  Prints cell-steps/s and the bytes the step moves through HBM (16 B per cell and step), and checks the
final state bit for bit against the staged schedule (halo=0) on the same input.  The tiled run is the plain drop-in
call: the stencil radius is measured by the probe kernel.
env ODEINTR_TILE_CPT = cells per thread (tile = 256 * CPT cells)."""
import ctypes as C, os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
logn = int(os.environ.get("LOGN", "28"))
steps = int(os.environ.get("STEPS", "100"))
check = os.environ.get("CHECK", "1") == "1"
cases = os.environ.get("CASES", "bistable,chain").split(",")
for name, src, pars, glob in (("bistable", BISTABLE_1D, BISTABLE_PARS, ""), ("chain", CHAIN, CHAIN_PARS, CHAIN_GLOBALS)):
    if name not in cases:
        continue
    n = 1 << logn
    if name == "chain":
        L = n // 2
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(L) / L), np.zeros(L)])
    else:
        x0 = (np.arange(n) % 7 < 3).astype(np.float64)
    fin = {}
    for halo in ((None, 0) if check else (None,)):
        s = odeintr_b200.compile_sys_openmp(name, src, pars, True, method="rk4", sys_dim=n, globals=glob, halo=halo).system
        tiled = halo is None
        if tiled:
            print(name, "interior kernel:", s.kernel_attributes("odeintr_lattice_rk4_tiled_interior"), flush=True)
        for rep in range(3 if tiled else 1):
            s.set_state(x0)
            rc = lib.odeintr_integrate_const(s._h, 0.0, steps * 0.01, 0.01, 1, 0)
            assert rc == 0, _abi.last_error()
            ms, st = s.last_kernel_ms(), s.last_steps()
            print("%s %s n=2^%d steps=%d %.2f ms -> %.4g cell-steps/s, %.1f GB/s at %d B/cell/step, launches %d" %
                  (name, "tiled(halo=%d)" % s.get_halo() if tiled else "staged", logn, st, ms, n * st / ms * 1e3,
                   (16.0 if tiled else 128.0) * n * st / ms / 1e6, 16 if tiled else 128, s.last_launches()), flush=True)
        fin[halo] = s.get_state().copy()
        s.close()
    if check:
        same = np.array_equal(fin[None].view(np.uint64), fin[0].view(np.uint64))
        print(name, "tiled == staged bit for bit:", same, flush=True)
        assert same
