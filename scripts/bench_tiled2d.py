"""On a B200: the reference's own large-system shape, a periodic M x M bistable lattice through laplace2D
(R/odeintr.R:393-413), M = 16384 (2^28 cells), rk4.  Tiled (odeintr_lattice_rk4_tiled2d, the default) against the
stage-per-launch kernels (halo=0), final states compared bit for bit."""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
M = int(os.environ.get("M", "16384"))
steps = int(os.environ.get("STEPS", "50"))
n = M * M
x0 = np.random.default_rng(1).integers(0, 2, n).astype(np.float64)
fin = {}
quick = os.environ.get("QUICK") == "1"  # tiled kernel only (parameter sweeps)
for name, src in ((("laplace2D", BISTABLE_2D),) if quick else (("laplace2D", BISTABLE_2D), ("laplace4", BISTABLE_DOC))):
    for halo in ((None,) if quick else (None, 0)):
        if name == "laplace4" and halo == 0:
            continue
        s = odeintr_b200.compile_sys_openmp("b2d", src, BISTABLE_PARS, True, method="rk4", sys_dim=n, halo=halo).system
        for rep in range(3 if halo is None else 1):
            s.set_state(x0)
            assert lib.odeintr_integrate_const(s._h, 0.0, steps * 0.01, 0.01, 1, 0) == 0, _abi.last_error()
            ms, st = s.last_kernel_ms(), s.last_steps()
            print("%s %dx%d %s steps=%d %.2f ms -> %.4g cell-steps/s, launches %d" %
                  (name, M, M, "tiled(halo=%d)" % s.get_halo() if s.get_halo() >= 0 else "staged", st, ms,
                   n * st / ms * 1e3, s.last_launches()), flush=True)
        if halo is None:
            print("  ", s.kernel_attributes("odeintr_lattice_rk4_tiled2d_h1"), flush=True)
        fin[(name, halo)] = s.get_state().copy()
        s.close()
if quick:
    sys.exit(0)
same = np.array_equal(fin[("laplace2D", None)].view(np.uint64), fin[("laplace2D", 0)].view(np.uint64))
same4 = np.array_equal(fin[("laplace4", None)].view(np.uint64), fin[("laplace2D", 0)].view(np.uint64))
print("tiled2d == staged bit for bit:", same, "| laplace4 form == laplace2D form:", same4, flush=True)
assert same and same4
