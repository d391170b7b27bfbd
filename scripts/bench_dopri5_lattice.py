"""On a B200: adaptive dopri5 (compile_sys_openmp's default method rk5_i, and rk5_a) on a 2^26-cell 1-D bistable
lattice: the temporally blocked attempt kernel (default) against the stage-per-launch path (ODEINTR_NO_TILED=1 in a
second process), final states written to gpurun_out for a bit comparison."""
import ctypes as C, os, sys, hashlib
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
logn = int(os.environ.get("LOGN", "26"))
n = 1 << logn
x0 = np.random.default_rng(1).uniform(0, 1, n)
tag = "staged" if os.environ.get("ODEINTR_NO_TILED") else ("tiled_block" if os.environ.get("ODEINTR_NO_WARP") else "tiled_warp")
for method in ("rk5_i", "rk5_a"):
    s = odeintr_b200.compile_sys_openmp("b1d", BISTABLE_1D, BISTABLE_PARS, True, method=method, sys_dim=n).system
    for rep in range(2):
        s.set_state(x0)
        at = np.array([0.0, 1.0, 5.0, 10.0])
        assert lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) == 0, _abi.last_error()
        ms = s.last_kernel_ms(); stt = s.stats()
        tries = int(stt["accepted"][0] + stt["rejected"][0])
        print("%s %s 2^%d bistable_at(0,1,5,10): %.2f ms, accepted %d rejected %d -> %.3f ms/attempt, %.4g cell-attempts/s, launches %d"
              % (tag, method, logn, ms, stt["accepted"][0], stt["rejected"][0], ms / tries, n * tries / ms * 1e3, s.last_launches()), flush=True)
    fin = s.get_state()
    print("%s %s final-state sha1 %s" % (tag, method, hashlib.sha1(fin.tobytes()).hexdigest()), flush=True)
    s.close()
