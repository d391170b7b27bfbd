#!/usr/bin/env python3
"""Regenerates tests/golden/oracle_golden.json from the oracle (oracle/build_oracle.py): small known-answer vectors
for every stepper family, stored as big-endian IEEE-754 hex so that drift of the oracle (or of the image's libm) is
caught by the CPU suite and the CUDA path is checked against committed numbers, not only against a live oracle run.
    python tests/golden/make_golden.py
(The Lorenz rk4 vectors of lorenz_rk4_golden.json come from SURVEY.md Appendix B and are not produced here.)"""
import json
import pathlib
import struct
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))
from oracle import build_oracle as bo  # noqa: E402
from systems import BISTABLE_1D, BISTABLE_PARS, ROBERTSON, ROBERTSON_PARS, STIFF, VDP  # noqa: E402


def hx(a):
    return [struct.pack(">d", float(v)).hex() for v in np.asarray(a, dtype=np.float64).reshape(-1)]


def cases():
    out = {}
    times = [0.0, 0.5, 2.0, 7.5, 20.0]
    for method in ("rk5_i", "rk5_a", "rk54_a", "rk78_a", "bs", "bsd"):
        o = bo.build(VDP, "mu", False, method, atol=1e-8, rtol=1e-8)
        r = o.integrate_times([2.0, 0.0], times, 0.01, pars=[1.5])
        out["vdp_at_" + method] = {"sys": "VDP", "mu": 1.5, "x0": [2.0, 0.0], "times": times, "dt": 0.01, "atol": 1e-8,
                                   "rtol": 1e-8, "rec": hx(r["rec"]), "accepted": r["accepted"], "rejected": r["rejected"]}
    for method in ("euler", "rk54", "rk5", "rk78"):
        o = bo.build(VDP, "mu", False, method)
        r = o.integrate_const([2.0, 0.0], 0.0, 2.0, 0.01, pars=[1.5], record=False)
        out["vdp_const_" + method] = {"sys": "VDP", "mu": 1.5, "x0": [2.0, 0.0], "t1": 2.0, "dt": 0.01, "x": hx(r["x"])}
    o = bo.build(STIFF, method="rosenbrock4")
    r = o.integrate_const([2.0, 1.0], 0.0, 5.0, 0.5)
    out["stiff_rosenbrock4"] = {"sys": "STIFF", "x0": [2.0, 1.0], "t1": 5.0, "dt": 0.5, "rec": hx(r["rec"]),
                                "accepted": r["accepted"], "rejected": r["rejected"]}
    o = bo.build(ROBERTSON, ROBERTSON_PARS, True, "rosenbrock4")
    tt = [1e-5, 1e-2, 1.0, 1e2, 1e5]
    adv = o.integrate_const([1.0, 0.0, 0.0], 0.0, tt[0], 1.0, record=False)
    r = o.integrate_times(adv["x"], tt, 1.0)
    out["robertson_at"] = {"sys": "ROBERTSON", "x0": [1.0, 0.0, 0.0], "times": tt, "rec": hx(r["rec"]),
                           "accepted": r["accepted"]}
    n = 64
    x0 = (np.arange(n) % 7 < 3).astype(np.float64)
    for method in ("rk4", "euler", "rk5_i"):
        o = bo.build(BISTABLE_1D, BISTABLE_PARS, True, method, sys_dim=n)
        r = o.integrate_adaptive(x0, 0.0, 3.0, 0.1)
        out["bistable64_" + method] = {"sys": "BISTABLE_1D", "n": n, "x0": "(arange(n) % 7 < 3)", "t1": 3.0, "dt": 0.1,
                                       "x": hx(r["x"])}
    return out


if __name__ == "__main__":
    (HERE / "oracle_golden.json").write_text(json.dumps(cases(), indent=1))
    print("wrote", HERE / "oracle_golden.json")
