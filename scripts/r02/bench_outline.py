#!/usr/bin/env python3
"""K2 / K3 with the step attempt outlined (one copy per kernel instead of one per integrate loop: ODEINTR_K2_OUTLINE /
ODEINTR_K3_OUTLINE) against the inlined default: Robertson rosenbrock4 and Van der Pol dopri5, 10 M instances."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
m = 10_000_000

def time_it(s, times, dt, X, reps=3):
    best = None
    for rep in range(reps):
        s.set_state(X)
        assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, dt) == 0, _abi.last_error()
        best = s.last_kernel_ms() if best is None else min(best, s.last_kernel_ms())
    st = s.stats()
    return best, int(st["accepted"].sum()), int(st["rejected"].sum()), s.get_state()

ref = {}
for outline in (False, True):
    for k in ("ODEINTR_K2_OUTLINE", "ODEINTR_K3_OUTLINE"):
        if outline: os.environ[k] = "1"
        else: os.environ.pop(k, None)
    for mb in (None, "4", "3"):
        if mb is None: os.environ.pop("ODEINTR_MIN_BLOCKS", None)
        else: os.environ["ODEINTR_MIN_BLOCKS"] = mb
        s = odeintr_b200.compile_implicit("rob", ROBERTSON, ROBERTSON_PARS).system
        lo = (C.c_double * 3)(0.02, 1e4, 3e7); hi = (C.c_double * 3)(0.08, 1e4, 3e7)
        assert lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi) == 0
        times = 10 ** np.linspace(-5, 5, 41)
        X = np.repeat(np.array([1.0, 0.0, 0.0])[:, None], m, axis=1)
        ms, acc, rej, fin = time_it(s, times, 1e-6, X)
        ref.setdefault("k3", fin)
        print(json.dumps({"kernel": "K3 robertson", "outlined": outline, "min_blocks": mb, "ms": ms, "accepted_per_s": acc / ms * 1e3,
                          "accepted": acc, "rejected": rej, "same_bits_as_default": bool(np.array_equal(fin.view(np.uint64), ref["k3"].view(np.uint64))),
                          **s.kernel_attributes("odeintr_adaptive_ensemble")}), flush=True)
        s.close()
        s = odeintr_b200.compile_sys("vpol", VDP, "mu", method="rk5_i", atol=1e-8, rtol=1e-8).system
        lo = (C.c_double * 1)(0.5); hi = (C.c_double * 1)(2.0)
        assert lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi) == 0
        times = np.arange(0, 20.0001, 0.2)
        X = np.repeat(np.array([2.0, 0.0])[:, None], m, axis=1)
        ms, acc, rej, fin = time_it(s, times, 0.01, X)
        ref.setdefault("k2", fin)
        print(json.dumps({"kernel": "K2 vdp", "outlined": outline, "min_blocks": mb, "ms": ms, "accepted_per_s": acc / ms * 1e3,
                          "accepted": acc, "rejected": rej, "same_bits_as_default": bool(np.array_equal(fin.view(np.uint64), ref["k2"].view(np.uint64))),
                          **s.kernel_attributes("odeintr_adaptive_ensemble")}), flush=True)
        s.close()
