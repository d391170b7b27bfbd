"""GPU parity for the fixed-step path (K1, configs 1 and 2 of BASELINE.json): the CUDA ensemble kernel,
called through the C-ABI, must reproduce the oracle bit for bit (north_star: rel 1e-12 for RK4 in FP64 —
for a chaotic system over 100 000 steps that is a bit-exactness criterion)."""
import ctypes as C
import json
import pathlib
import struct
import warnings

import numpy as np
import pytest

import odeintr_b200
from odeintr_b200 import _abi
from oracle import build_oracle as bo
from philox import uniform_state
from systems import LOGISTIC, LORENZ, LORENZ_PARS

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((pathlib.Path(__file__).parent / "golden" / "lorenz_rk4_golden.json").read_text())


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def hexd(v):
    return struct.pack(">d", float(v)).hex()


@pytest.fixture(scope="module")
def lorenz_const():
    env = {}
    code = odeintr_b200.compile_sys("lorenz", LORENZ, LORENZ_PARS, True, method="rk4", env=env)
    return env, code.system, bo.build(LORENZ, LORENZ_PARS, True, "rk4")


@pytest.fixture(scope="module")
def lorenz_dyn():
    env = {}
    code = odeintr_b200.compile_sys("lorenz2", LORENZ, LORENZ_PARS, False, method="rk4", env=env)
    return env, code.system, bo.build(LORENZ, LORENZ_PARS, False, "rk4")


def test_config1_single_trajectory_bit_exact(lorenz_const):
    # BASELINE.json configs[0]: lorenz(rep(1,3), 100, 0.001) with rk4
    env, system, ora = lorenz_const
    df = env["lorenz"](np.ones(3), 100, 0.001)
    ref = ora.integrate_const(np.ones(3), 0, 100, 0.001)
    assert list(df.columns) == ["Time", "X1", "X2", "X3"] and len(df) == 100001
    got = df[["X1", "X2", "X3"]].to_numpy()
    assert np.array_equal(bits(got), bits(ref["rec"]))
    assert np.array_equal(bits(df["Time"].to_numpy()), bits(ref["t"]))
    # relative error criterion of north_star, stated explicitly (bit-exact implies it): 1e-12
    assert np.max(np.abs(got - ref["rec"]) / np.maximum(np.abs(ref["rec"]), 1e-300)) <= 1e-12
    case = GOLDEN["cases"][0]
    for n, want in case["steps"].items():
        assert [hexd(v) for v in got[int(n)]] == want
    assert np.array_equal(env["lorenz_get_state"](), got[-1])


def test_runtime_parameters_exact_double(lorenz_dyn):
    env, system, ora = lorenz_dyn
    assert env["lorenz2_get_params"]() == {"sigma": 10.0, "R": 28.0, "b": 2.66666666666667}
    df = env["lorenz2"](np.ones(3), 10, 0.001)  # defaults: the 15-digit text value of b
    assert [hexd(v) for v in df.iloc[10000, 1:]] == GOLDEN["cases"][0]["steps"]["10000"]
    env["lorenz2_set_params"](b=8 / 3)
    df = env["lorenz2"](np.ones(3), 10, 0.001)
    assert [hexd(v) for v in df.iloc[10000, 1:]] == GOLDEN["cases"][1]["steps"]["10000"]
    env["lorenz2_set_params"]()  # R default arguments restore the declared values


def test_config2_miniature_ensemble_philox(lorenz_const):
    # BASELINE.json configs[1] at 65 521 instances (prime: exercises the ragged last block)
    env, system, ora = lorenz_const
    lib = _abi.load()
    m, seed = 65521, 20260101
    lo = np.array([-20.0, -25.0, 5.0])
    hi = np.array([20.0, 25.0, 45.0])
    rc = lib.odeintr_fill_state_uniform(system._h, m, seed, 1000, lo.ctypes.data_as(_abi.c_dbl_p),
                                        hi.ctypes.data_as(_abi.c_dbl_p))
    assert rc == 0, _abi.last_error()
    x0 = system.get_state()
    want = uniform_state(np.arange(1000, 1000 + m), 3, seed, lo, hi)
    assert np.array_equal(bits(x0), bits(want))
    assert x0[0].min() >= -20 and x0[0].max() < 20 and abs(x0[2].mean() - 25) < 0.5
    rc = lib.odeintr_integrate_const(system._h, 0.0, 1.0, 0.001, 100, 1)
    assert rc == 0, _abi.last_error()
    rec = system.get_output()
    assert rec.x.shape == (11, 3, m)
    assert np.array_equal(bits(rec.time), bits(np.array([0.0 + (100 * s) * 0.001 for s in range(11)])))
    e = ora.ensemble(0, x0.copy(), None, 0.0, 1.0, 0.001, obs_stride=100, out_rows=11, threads=ora.max_threads())
    assert np.array_equal(bits(rec.x), bits(e["out"]))
    assert np.array_equal(bits(system.get_state()), bits(e["x"]))
    # get_output for an instance range = data-frame columns per instance
    sub = np.empty((5, 3, 11))
    rc = lib.odeintr_get_output(system._h, sub.ctypes.data_as(_abi.c_dbl_p), sub.size, 777, 5)
    assert rc == 0, _abi.last_error()
    assert np.array_equal(sub, np.transpose(rec.x[:, :, 777:782], (2, 1, 0)))


def test_parameter_sweep_per_instance(lorenz_dyn):
    env, system, ora = lorenz_dyn
    m = 1000
    Rs = np.linspace(20, 35, m)
    env["lorenz2_set_params"](R=Rs, b=8 / 3)
    rec = env["lorenz2"](np.array([1.0, 2.0, 3.0]), 2.0, 0.01)  # one state, replicated over the sweep
    P = np.stack([np.full(m, 10.0), Rs, np.full(m, 8 / 3)])
    X = np.repeat(np.array([[1.0], [2.0], [3.0]]), m, axis=1)
    e = ora.ensemble(0, X, P, 0.0, 2.0, 0.01, out_rows=rec.rows, threads=ora.max_threads())
    assert rec.x.shape == (201, 3, m) and np.array_equal(bits(rec.x), bits(e["out"]))
    env["lorenz2_set_params"]()


@pytest.mark.parametrize("method", ["euler", "rk4", "rk54", "rk5", "rk78"])
@pytest.mark.parametrize("flags", [0, _abi.FLAG_KEEP_ZERO_TERMS])
def test_every_fixed_step_method_bit_exact(method, flags):
    sys_t = "dxdt[0] = x[1] + 0.1 * t; dxdt[1] = -k * x[0] - 0.05 * x[1] * x[1] * x[1] + std::sqrt(t + 1.0) / (2.0 + x[0] * x[0])"
    env = {}
    odeintr_b200.compile_sys("osc", sys_t, {"k": 2.5}, method=method, env=env, flags=flags)
    ora = bo.build(sys_t, {"k": 2.5}, False, method)
    x0 = np.array([0.3, -1.2])
    df = env["osc"](x0, 7.0, 0.013, start=0.5)
    ref = ora.integrate_const(x0, 0.5, 7.5, 0.013, pars=[2.5])
    assert len(df) == ref["rows"]
    assert np.array_equal(bits(df[["X1", "X2"]].to_numpy()), bits(ref["rec"]))
    assert np.array_equal(bits(df["Time"].to_numpy()), bits(ref["t"]))
    # name_at: integrate_const advance + integrate_times with remainder steps
    times = np.array([1.0, 1.05, 1.3, 1.3000000001, 2.77, 5.0])
    df = env["osc_at"](x0, times, 0.1, start=0.25)
    adv = ora.integrate_const(x0, 0.25, times[0], 0.1, pars=[2.5], record=False)
    ref = ora.integrate_times(adv["x"], times, 0.1, pars=[2.5])
    assert np.array_equal(bits(df[["X1", "X2"]].to_numpy()), bits(ref["rec"]))
    assert np.array_equal(df["Time"].to_numpy(), times)
    # name_continue_at resumes from the last recorded time with the retained state
    more = np.array([5.5, 6.0, 6.25])
    df2 = env["osc_continue_at"](more, 0.1)
    adv2 = ora.integrate_const(ref["x"], times[-1], more[0], 0.1, pars=[2.5], record=False)
    ref2 = ora.integrate_times(adv2["x"], more, 0.1, pars=[2.5])
    assert np.array_equal(bits(df2[["X1", "X2"]].to_numpy()), bits(ref2["rec"]))
    # name_adap / name_no_record: integrate_adaptive for plain steppers = n steps + one step onto t1
    df = env["osc_adap"](x0, 3.3, 0.25)
    ref = ora.integrate_adaptive(x0, 0.0, 3.3, 0.25, pars=[2.5], record=True)
    assert np.array_equal(bits(df[["X1", "X2"]].to_numpy()), bits(ref["rec"]))
    assert np.array_equal(bits(df["Time"].to_numpy()), bits(ref["t"])) and df["Time"].iloc[-1] == 3.3
    final = env["osc_no_record"](x0, 3.3, 0.25)
    assert np.array_equal(bits(final), bits(ref["x"]))
    assert np.array_equal(env["osc_get_state"](), final)


def test_step_guard_rows_and_errors():
    # tests/testthat/test_odeintr.R:40-46 — step 100 on duration 40 -> warning, 4 rows, first row == init
    env = {}
    odeintr_b200.compile_sys("logi", LOGISTIC, method="rk4", env=env)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        df = env["logi"](0.01, 40, 100)
    assert any("Step size too large -- adjusting" in str(x.message) for x in w)
    assert df.shape == (4, 2) and df.iloc[0, 1] == 0.01
    df = env["logi"](0.01, 40)
    assert df.shape == (41, 2) and df.iloc[0, 1] == 0.01
    with pytest.raises(odeintr_b200.OdeintrError, match="Invalid initial state"):
        env["logi"]([0.1, 0.2], 40)
    env["logi_reset_observer"]()
    with pytest.raises(odeintr_b200.OdeintrError):
        env["logi_continue_at"]([1.0])  # the reference reads rec_t.back() of an empty record here


def test_host_streaming_path_matches_resident_path(lorenz_dyn):
    # odeintr_integrate_const_host: chunked, overlapped copies; ragged last chunk; per-instance parameters
    env, system, ora = lorenz_dyn
    rng = np.random.default_rng(3)
    m = 10_000
    x0 = rng.uniform(-10, 10, (3, m))
    P = np.stack([np.full(m, 10.0), rng.uniform(20, 35, m), np.full(m, 8 / 3)])
    rec_h, xf_h = system.run_host(x0, 1.0, 0.01, obs_stride=7, pars=P, chunk=1536)
    e = ora.ensemble(0, x0.copy(), P, 0.0, 1.0, 0.01, obs_stride=7, out_rows=rec_h.rows, threads=ora.max_threads())
    assert rec_h.x.shape == (16, 3, m)
    assert np.array_equal(bits(rec_h.x), bits(e["out"])) and np.array_equal(bits(xf_h), bits(e["x"]))
    assert np.array_equal(bits(rec_h.time), bits(np.array([(7 * s) * 0.01 for s in range(15)] + [100 * 0.01])))
    # broadcast parameters, no observer
    _, xf = system.run_host(x0, 1.0, 0.01, pars=np.array([[10.0], [28.0], [8 / 3]]), chunk=4096, record=False)
    e = ora.ensemble(0, x0.copy(), np.array([10.0, 28.0, 8 / 3]), 0.0, 1.0, 0.01, threads=ora.max_threads())
    assert np.array_equal(bits(xf), bits(e["x"]))


def test_fast_mode_is_close_but_not_the_parity_path(lorenz_const):
    env_f = {}
    odeintr_b200.compile_sys("lorenz_f", LORENZ, LORENZ_PARS, True, method="rk4", env=env_f, flags=_abi.FLAG_FMAD)
    a = env_f["lorenz_f"](np.ones(3), 1.0, 0.001)[["X1", "X2", "X3"]].to_numpy()
    b = lorenz_const[0]["lorenz"](np.ones(3), 1.0, 0.001)[["X1", "X2", "X3"]].to_numpy()
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-10  # FMA contraction: close on a short horizon only


def test_full_size_ensemble_sampled_against_oracle(lorenz_const):
    # BASELINE.json configs[1] at full size (100 M instances, T = 1, observer every 0.1): 4 096 instances
    # drawn from all over the index range are re-run by the oracle from their regenerated Philox states.
    env, system, ora = lorenz_const
    lib = _abi.load()
    m, seed = 100_000_000, 20260101
    lo = np.array([-20.0, -25.0, 5.0])
    hi = np.array([20.0, 25.0, 45.0])
    rc = lib.odeintr_fill_state_uniform(system._h, m, seed, 0, lo.ctypes.data_as(_abi.c_dbl_p), hi.ctypes.data_as(_abi.c_dbl_p))
    assert rc == 0, _abi.last_error()
    rc = lib.odeintr_integrate_const(system._h, 0.0, 1.0, 0.001, 100, 1)
    assert rc == 0, _abi.last_error()
    assert lib.odeintr_output_rows(system._h) == 11 and system.last_steps() == 1000
    rng = np.random.default_rng(7)
    starts = np.concatenate([[0, m - 64], rng.integers(0, m - 64, 62)])
    for s0 in starts:
        sub = np.empty((64, 3, 11))
        rc = lib.odeintr_get_output(system._h, sub.ctypes.data_as(_abi.c_dbl_p), sub.size, int(s0), 64)
        assert rc == 0, _abi.last_error()
        x0 = uniform_state(np.arange(s0, s0 + 64), 3, seed, lo, hi)
        assert np.array_equal(bits(sub[:, :, 0]), bits(x0.T))
        e = ora.ensemble(0, x0.copy(), None, 0.0, 1.0, 0.001, obs_stride=100, out_rows=11, threads=1)
        assert np.array_equal(bits(sub), bits(np.transpose(e["out"], (2, 1, 0))))
    assert lib.odeintr_release_record(system._h) == 0  # give the 26 GB record back
    system.set_state(np.ones(3))
