"""Scratch exploration on a B200: probes + Lorenz rk4 ensemble throughput vs ensemble size / block size."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odeintr_b200
from odeintr_b200 import _abi

lib = _abi.load()
LOR = """
  dxdt[0] = sigma * (x[1] - x[0]);
  dxdt[1] = R * x[0] - x[1] - x[0] * x[2];
  dxdt[2] = -b * x[2] + x[0] * x[1];
"""
PARS = {"sigma": 10, "R": 28, "b": 8 / 3}
os.system("nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv")
for kind in (0, 1):
    v, ms = C.c_double(), C.c_double()
    rc = lib.odeintr_fp64_probe(kind, 0, C.byref(v), C.byref(ms))
    print("fp64 probe kind", kind, "rc", rc, "DP instr/s %.4g" % v.value, "ms %.3f" % ms.value, flush=True)
bw = C.c_double()
lib.odeintr_hbm_probe(0, 4 << 30, C.byref(bw))
print("hbm copy probe GB/s %.1f" % (bw.value / 1e9), flush=True)

flagsets = [("exact", 0), ("fmad", 1), ("zero_terms", 2)]
for label, flags in flagsets:
    code = odeintr_b200.compile_sys("lorenz", LOR, PARS, True, method="rk4", flags=flags)
    s = code.system
    print(label, s.kernel_attributes("odeintr_plain_ensemble"), flush=True)
    lo = (C.c_double * 3)(-20, -25, 5)
    hi = (C.c_double * 3)(20, 25, 45)
    for m in (1 << 20, 1 << 24, 100_000_000):
        rc = lib.odeintr_fill_state_uniform(s._h, m, 20260101, 0, lo, hi)
        assert rc == 0, _abi.last_error()
        for rep in range(3):
            if rep:
                lib.odeintr_fill_state_uniform(s._h, m, 20260101, 0, lo, hi)
            t0 = time.time()
            rc = lib.odeintr_integrate_const(s._h, 0.0, 1.0, 0.001, 100, 1)
            assert rc == 0, _abi.last_error()
            ms = s.last_kernel_ms()
            wall = time.time() - t0
        steps = s.last_steps()
        print("%s m=%d steps=%d rows=%d kernel %.3f ms wall %.3f s -> %.4g traj-steps/s" %
              (label, m, steps, lib.odeintr_output_rows(s._h), ms, wall, m * steps / (ms * 1e-3)), flush=True)
        # no observer
        lib.odeintr_fill_state_uniform(s._h, m, 20260101, 0, lo, hi)
        rc = lib.odeintr_integrate_const(s._h, 0.0, 1.0, 0.001, 1, 0)
        ms = s.last_kernel_ms()
        print("%s m=%d no_record kernel %.3f ms -> %.4g traj-steps/s" % (label, m, ms, m * steps / (ms * 1e-3)), flush=True)
    s.close()
