"""On a B200: the reference's documented large-system example on its default method -- bistable_at(inic, times) with
rk5_i on an M x M lattice (R/odeintr.R:393-421): the streaming attempt kernel K4ds2 (default) against the seven
stage-per-launch kernels (ODEINTR_NO_STREAM2D=1 in a second process); final-state sha1 for the bit comparison."""
import ctypes as C, os, sys, hashlib
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
m = int(os.environ.get("M", "8192"))
n = m * m
x0 = np.random.default_rng(1).uniform(0, 1, n)
tag = "staged" if os.environ.get("ODEINTR_NO_STREAM2D") else "stream"
for method in os.environ.get("METHODS", "rk5_i,rk5_a").split(","):
    s = odeintr_b200.compile_sys_openmp("b2d", BISTABLE_2D, BISTABLE_PARS, True, method=method, sys_dim=n).system
    for rep in range(2):
        s.set_state(x0)
        at = np.array([0.0, 1.0, 5.0, 10.0])
        assert lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) == 0, _abi.last_error()
        ms = s.last_kernel_ms(); stt = s.stats()
        tries = int(stt["accepted"][0] + stt["rejected"][0])
        print("%s %s %dx%d bistable_at(0,1,5,10): %.2f ms, accepted %d rejected %d -> %.3f ms/attempt, %.4g cell-attempts/s, launches %d"
              % (tag, method, m, m, ms, stt["accepted"][0], stt["rejected"][0], ms / tries, n * tries / ms * 1e3, s.last_launches()), flush=True)
    fin = s.get_state()
    print("%s %s final-state sha1 %s" % (tag, method, hashlib.sha1(fin.tobytes()).hexdigest()), flush=True)
    s.close()
