#!/usr/bin/env python3
"""K6 measurement: heterogeneous adaptive ensembles (the sweeps of scripts/explore_divergence.py, 1 Mi instances) in
index order, in cost-proxy order (pilot + sort inside the timed call) and -- for the parameter sweep -- physically
sorted.  Per-instance results are compared bit for bit between the orders."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import *
lib = _abi.load()
m = 1 << 20

def run(s, times, dt, X=None, fill=None, reps=3):
    best = None
    for rep in range(reps):
        if X is not None: s.set_state(X)
        else: fill()
        assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, dt) == 0, _abi.last_error()
        best = s.last_kernel_ms() if best is None else min(best, s.last_kernel_ms())
    st = s.stats()
    return best, st["accepted"], st["rejected"], s.get_state()

def line(label, ms, acc, rej, **kw):
    tries = int(acc.sum() + rej.sum())
    w = (acc + rej).reshape(-1, 32)
    d = {"case": label, "ms": round(ms, 3), "tries_per_s": tries / ms * 1e3, "warp_efficiency_by_tries": float((acc + rej).sum() / (32.0 * w.max(axis=1).sum()))}
    d.update(kw)
    print(json.dumps(d), flush=True)

s = odeintr_b200.compile_sys("vpol", VDP, "mu", method="rk5_i", atol=1e-8, rtol=1e-8).system
mus = np.random.default_rng(0).uniform(0.1, 10.0, m)
times = np.arange(0, 20.0001, 0.2)
X = np.repeat(np.array([2.0, 0.0])[:, None], m, axis=1)
s.set_params(mu=mus)
ms_i, acc_i, rej_i, fin_i = run(s, times, 0.01, X)
line("vdp shuffled mu, index order", ms_i, acc_i, rej_i)
s.set_instance_order("cost")
ms_c, acc_c, rej_c, fin_c = run(s, times, 0.01, X)
same = bool(np.array_equal(fin_i.view(np.uint64), fin_c.view(np.uint64)) and np.array_equal(acc_i, acc_c) and np.array_equal(rej_i, rej_c))
line("vdp shuffled mu, cost-proxy order (pilot + sort in the timed call)", ms_c, acc_c, rej_c, speedup=ms_i / ms_c, identical=same)
s.set_instance_order("index")
s.set_params(mu=np.sort(mus))
ms_s, acc_s, rej_s, _ = run(s, times, 0.01, X)
line("vdp physically sorted mu, index order", ms_s, acc_s, rej_s)
s.close()

s = odeintr_b200.compile_sys("lor", LORENZ, LORENZ_PARS, True, method="rk5_i", atol=1e-8, rtol=1e-8).system
lo = (C.c_double * 3)(-20, -25, 5); hi = (C.c_double * 3)(20, 25, 45)
times = np.arange(0, 10.0001, 0.1)
fill = lambda: lib.odeintr_fill_state_uniform(s._h, m, 1, 0, lo, hi)
ms_i, acc_i, rej_i, fin_i = run(s, times, 0.01, fill=fill)
line("lorenz random ICs, index order", ms_i, acc_i, rej_i)
s.set_instance_order("cost")
ms_c, acc_c, rej_c, fin_c = run(s, times, 0.01, fill=fill)
same = bool(np.array_equal(fin_i.view(np.uint64), fin_c.view(np.uint64)) and np.array_equal(acc_i, acc_c))
line("lorenz random ICs, cost-proxy order", ms_c, acc_c, rej_c, speedup=ms_i / ms_c, identical=same)
s.close()

s = odeintr_b200.compile_implicit("rob", ROBERTSON, ROBERTSON_PARS).system
al = np.random.default_rng(1).uniform(0.02, 0.08, m)
s.set_params(alpha=al, beta=1e4, gamma=3e7)
times = 10 ** np.linspace(-5, 5, 41)
X = np.repeat(np.array([1.0, 0.0, 0.0])[:, None], m, axis=1)
ms_i, acc_i, rej_i, fin_i = run(s, times, 1e-6, X)
line("robertson shuffled alpha, index order", ms_i, acc_i, rej_i)
s.set_instance_order("cost")
ms_c, acc_c, rej_c, fin_c = run(s, times, 1e-6, X)
line("robertson shuffled alpha, cost-proxy order", ms_c, acc_c, rej_c, speedup=ms_i / ms_c,
     identical=bool(np.array_equal(fin_i.view(np.uint64), fin_c.view(np.uint64))))
s.close()
