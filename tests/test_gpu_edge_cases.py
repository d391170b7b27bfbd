"""Edge cases of the generated-function surface on the GPU, each against the oracle: backward integration,
degenerate time specifications, vector parameters and helper functions in `globals`, a wider system."""
import numpy as np
import pytest

import odeintr_b200
from oracle import build_oracle as bo
from test_gpu_adaptive import close

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


OSC = "dxdt[0] = x[1]; dxdt[1] = -pars[0] * x[0] - damp(x[1])"
OSC_GLOBALS = "double damp(double v) { return pars[1] * v * v * v; }"


def test_vector_parameters_and_globals_helper_ensemble():
    # pars = 2 -> `pars[k]` vector (R/utils.R:209-216) used from a helper defined in `globals`
    env = {}
    odeintr_b200.compile_sys("osc", OSC, pars=2, method="rk4", globals=OSC_GLOBALS, env=env)
    ora = bo.build(OSC, 2, False, "rk4", globals=OSC_GLOBALS)
    m = 33
    P = np.stack([np.linspace(1, 4, m), np.linspace(0, 0.3, m)])
    env["osc_set_params"](P)
    assert np.array_equal(env["osc_get_params"](), P)
    x0 = np.stack([np.linspace(-1, 1, m), np.ones(m)])
    rec = env["osc"](x0, 3.0, 0.01)
    e = ora.ensemble(0, x0.copy(), P, 0.0, 3.0, 0.01, out_rows=rec.rows, threads=2)
    assert np.array_equal(bits(rec.x), bits(e["out"]))
    with pytest.raises(odeintr_b200.OdeintrError, match="Invalid parameter vector"):
        env["osc_set_params"](np.ones(3))


@pytest.mark.parametrize("method", ["rk4", "rk5_i", "rk5_a"])
def test_backward_integration_negative_step(method):
    # Boost's loops are sign-aware (less_with_sign): duration < 0 with step_size < 0 integrates backwards
    sys_t = "dxdt[0] = x[1]; dxdt[1] = -x[0] + 0.1 * (1 - x[0] * x[0]) * x[1]"
    env = {}
    odeintr_b200.compile_sys("b", sys_t, method=method, env=env)
    ora = bo.build(sys_t, method=method)
    x0 = np.array([1.0, 0.5])
    df = env["b"](x0, -2.0, -0.05)
    ref = ora.integrate_const(x0, 0.0, -2.0, -0.05)
    assert len(df) == ref["rows"] and np.array_equal(df["Time"].to_numpy(), ref["t"]) and df["Time"].iloc[-1] < -1.9
    got = df[["X1", "X2"]].to_numpy()
    if method == "rk4":
        assert np.array_equal(bits(got), bits(ref["rec"]))
    else:
        close(got, ref["rec"], 1e-6)
    times = np.array([0.0, -0.3, -1.0, -1.5])
    df = env["b_at"](x0, times, -0.1)
    ref = ora.integrate_times(x0, times, -0.1)
    close(df[["X1", "X2"]].to_numpy(), ref["rec"], 1e-6)


@pytest.mark.parametrize("method", ["rk4", "rk5_i"])
def test_degenerate_time_specifications(method):
    env = {}
    odeintr_b200.compile_sys("d", "dxdt = -x", method=method, env=env)
    ora = bo.build("dxdt = -x", method=method)
    # a single requested time: one row, state untouched
    df = env["d_at"](0.7, np.array([2.0]), 0.1, start=2.0)
    assert df.shape == (1, 2) and df.iloc[0, 0] == 2.0 and df.iloc[0, 1] == 0.7
    # repeated times produce repeated rows
    times = np.array([0.0, 0.5, 0.5, 1.0])
    df = env["d_at"](0.7, times, 0.1)
    ref = ora.integrate_times([0.7], times, 0.1)
    assert len(df) == 4
    close(df[["X1"]].to_numpy(), ref["rec"], 1e-6)
    # start past times[0]: the advance phase does nothing (integrate_const takes no step), SURVEY.md §3.3
    df = env["d_at"](0.7, np.array([1.0, 2.0]), 5.0, start=0.0)
    adv = ora.integrate_const([0.7], 0.0, 1.0, 5.0, record=False)
    ref = ora.integrate_times(adv["x"], [1.0, 2.0], 5.0)
    close(df[["X1"]].to_numpy(), ref["rec"], 1e-6)
    assert df.iloc[0, 1] == 0.7
    # no_record with zero duration returns the initial state
    assert env["d_no_record"](0.7, 0.0, 0.1)[0] == 0.7


def test_wider_system_lorenz96_in_registers():
    # N = 8 per thread (rk4: 6 vectors x 8 doubles): still register resident, bit-exact
    n = 8
    lines = ["dxdt[%d] = (x[%d] - x[%d]) * x[%d] - x[%d] + F" % (i, (i + 1) % n, (i - 2) % n, (i - 1) % n, i)
             for i in range(n)]
    src = "; ".join(lines)
    env = {}
    code = odeintr_b200.compile_sys("l96", src, {"F": 8.0}, method="rk4", env=env)
    ora = bo.build(src, {"F": 8.0}, False, "rk4")
    rng = np.random.default_rng(61)
    m = 500
    x0 = 8.0 + rng.normal(0, 0.5, (n, m))
    Fs = np.linspace(6, 10, m)
    env["l96_set_params"](F=Fs)
    rec = env["l96"](x0, 1.0, 0.01)
    e = ora.ensemble(0, x0.copy(), Fs[None, :], 0.0, 1.0, 0.01, out_rows=rec.rows, threads=2)
    assert np.array_equal(bits(rec.x), bits(e["out"]))
    assert code.system.kernel_attributes("odeintr_plain_ensemble")["local_bytes"] == 0


@pytest.mark.parametrize("kind", ["rk5_i", "implicit"])
def test_at_with_first_time_one_step_after_start(kind):
    # name_at(init, times, step_size, start) with times[0] - start == step_size: Boost's dense integrate_const
    # evaluates its interpolant before any step was taken (undefined behaviour in the reference: unset stage
    # vectors).  Product and oracle make the same conscious fix: real steps to times[0].
    env = {}
    src = "dxdt[0] = -x[0] * x[0]"
    if kind == "implicit":
        odeintr_b200.compile_implicit("q", src, env=env)
        ora = bo.build(src, method="rosenbrock4")
    else:
        odeintr_b200.compile_sys("q", src, env=env)
        ora = bo.build(src, method="rk5_i")
    times = np.array([1.0, 10.0, 100.0])
    df = env["q_at"](2.0, times)  # default step_size = 1, start = 0
    adv = ora.integrate_const([2.0], 0.0, 1.0, 1.0, record=False)
    ref = ora.integrate_times(adv["x"], times, 1.0)
    assert np.all(np.isfinite(df["X1"]))
    close(df[["X1"]].to_numpy(), ref["rec"], 1e-6)
    exact = 2.0 / (1.0 + 2.0 * times)
    assert np.max(np.abs(df["X1"].to_numpy() - exact)) < 1e-4


def test_custom_observer_is_replayed_over_recorded_samples():
    # README.md:66-90 style: compile_sys(..., observer = function(x, t) c(Prey = x[1], Predator = x[2], Ratio = x[1] / x[2]))
    lv = "dxdt[0] = x[0] - x[0] * x[1]; dxdt[1] = x[0] * x[1] - x[1]"
    env = {}
    obs = lambda x, t: {"Prey": x[0], "Predator": x[1], "Ratio": x[0] / x[1]}
    odeintr_b200.compile_sys("lv", lv, method="rk4", observer=obs, env=env)
    plain = {}
    odeintr_b200.compile_sys("lv0", lv, method="rk4", env=plain)
    df = env["lv"](np.array([1.0, 0.5]), 5.0, 0.1)
    ref = plain["lv0"](np.array([1.0, 0.5]), 5.0, 0.1)
    assert list(df.columns) == ["Time", "Prey", "Predator", "Ratio"] and len(df) == len(ref)
    assert np.array_equal(df["Prey"].to_numpy(), ref["X1"].to_numpy())
    assert np.array_equal(df["Ratio"].to_numpy(), (ref["X1"] / ref["X2"]).to_numpy())
    # unnamed results get X1..Xk (proc_output), samples for which the observer returns nothing are skipped
    env["lv_set_observer"](lambda x, t: [x[0] + x[1]] if t >= 1.0 else [])
    out = env["lv_get_output"]()
    assert list(out.columns) == ["Time", "X1"] and out["Time"].iloc[0] >= 1.0 - 1e-12
    assert env["lv_get_observer"]() is not None
    env["lv_set_output_processor"](lambda res: len(res[0]))
    assert env["lv_get_output"]() == len(out)
