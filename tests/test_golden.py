"""Committed known-answer vectors (tests/golden/oracle_golden.json, made by tests/golden/make_golden.py).
CPU: the oracle must still reproduce them bit for bit (catches drift of the restatement or of libm).
GPU: the CUDA path is compared with the committed numbers — bit-exact for fixed-step steppers and the large-system
paths, within 10 x rtol for the adaptive ensemble kernels (whose controller pow() is CUDA's, not glibc's)."""
import json
import pathlib
import struct

import numpy as np
import pytest

from systems import BISTABLE_1D, BISTABLE_PARS, ROBERTSON, ROBERTSON_PARS, STIFF, VDP

GOLD = json.loads((pathlib.Path(__file__).parent / "golden" / "oracle_golden.json").read_text())


def unhex(h):
    return np.array([struct.unpack(">d", bytes.fromhex(v))[0] for v in h])


def test_oracle_reproduces_committed_vectors():
    import sys

    sys.path.insert(0, str(pathlib.Path(__file__).parent / "golden"))
    import make_golden

    fresh = make_golden.cases()
    assert sorted(fresh) == sorted(GOLD)
    for name, case in GOLD.items():
        for key in ("rec", "x"):
            if key in case:
                assert fresh[name][key] == case[key], name
        for key in ("accepted", "rejected"):
            if key in case:
                assert fresh[name][key] == case[key], name


@pytest.mark.gpu
def test_cuda_path_against_committed_vectors():
    import odeintr_b200

    def tol(got, want, rtol):
        return np.max(np.abs(got - want) / (1 + np.abs(want))) <= 10 * rtol

    for method in ("rk5_i", "rk5_a", "rk54_a", "rk78_a", "bs", "bsd"):
        c = GOLD["vdp_at_" + method]
        env = {}
        code = odeintr_b200.compile_sys("v", VDP, "mu", method=method, atol=c["atol"], rtol=c["rtol"], env=env)
        env["v_set_params"](mu=c["mu"])
        df = env["v_at"](np.array(c["x0"]), np.array(c["times"]), c["dt"])
        eff = 1e-6 if method in ("bs", "bsd") else c["rtol"]
        assert tol(df[["X1", "X2"]].to_numpy().reshape(-1), unhex(c["rec"]), eff), method
        st = code.system.stats()
        assert abs(int(st["accepted"][0]) - c["accepted"]) <= (2 if method in ("bs", "bsd") else 0), method
    for method in ("euler", "rk54", "rk5", "rk78"):
        c = GOLD["vdp_const_" + method]
        env = {}
        odeintr_b200.compile_sys("v", VDP, "mu", method=method, env=env)
        env["v_set_params"](mu=c["mu"])
        df = env["v"](np.array(c["x0"]), c["t1"], c["dt"])  # integrate_const, like the golden run
        x = np.ascontiguousarray(df[["X1", "X2"]].to_numpy()[-1])
        assert np.array_equal(x.view(np.uint64), unhex(c["x"]).view(np.uint64)), method
    c = GOLD["stiff_rosenbrock4"]
    env = {}
    code = odeintr_b200.compile_implicit("s", STIFF, env=env)
    df = env["s"](np.array(c["x0"]), c["t1"], c["dt"])
    assert tol(df[["X1", "X2"]].to_numpy().reshape(-1), unhex(c["rec"]), 1e-6)
    assert int(code.system.stats()["accepted"][0]) == c["accepted"]
    c = GOLD["robertson_at"]
    env = {}
    odeintr_b200.compile_implicit("r", ROBERTSON, ROBERTSON_PARS, True, env=env)
    df = env["r_at"](np.array(c["x0"]), np.array(c["times"]))
    assert tol(df[["X1", "X2", "X3"]].to_numpy().reshape(-1), unhex(c["rec"]), 1e-6)
    n = 64
    x0 = (np.arange(n) % 7 < 3).astype(np.float64)
    for method in ("rk4", "euler", "rk5_i"):
        c = GOLD["bistable64_" + method]
        env = {}
        odeintr_b200.compile_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, method=method, sys_dim=n, env=env)
        x = env["b_no_record"](x0, c["t1"], c["dt"])
        assert np.array_equal(x.view(np.uint64), unhex(c["x"]).view(np.uint64)), method
