#!/usr/bin/env python3
"""K2 (config 3: Van der Pol dopri5 ensemble, 10 M instances) with the tableau in constant memory (default) against
literals (ODEINTR_DP5_LITERALS=1)."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
m = 10_000_000
ref = None
for lit in (False, True, False):
    if lit: os.environ["ODEINTR_DP5_LITERALS"] = "1"
    else: os.environ.pop("ODEINTR_DP5_LITERALS", None)
    s = odeintr_b200.compile_sys("vpol", VDP, "mu", method="rk5_i", atol=1e-8, rtol=1e-8).system
    lo = (C.c_double * 1)(0.5); hi = (C.c_double * 1)(2.0)
    assert lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi) == 0
    times = np.arange(0, 20.0001, 0.2)
    X = np.repeat(np.array([2.0, 0.0])[:, None], m, axis=1)
    best = None
    for rep in range(3):
        s.set_state(X)
        assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, 0.01) == 0, _abi.last_error()
        best = s.last_kernel_ms() if best is None else min(best, s.last_kernel_ms())
    st = s.stats(); fin = s.get_state()
    if ref is None: ref = fin
    print(json.dumps({"kernel": "K2 vdp", "tableau": "literals" if lit else "constant memory", "ms": best,
                      "accepted_per_s": int(st["accepted"].sum()) / best * 1e3,
                      "same_bits": bool(np.array_equal(fin.view(np.uint64), ref.view(np.uint64))),
                      **s.kernel_attributes("odeintr_adaptive_ensemble")}), flush=True)
    s.close()
