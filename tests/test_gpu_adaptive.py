"""GPU parity for the adaptive explicit steppers (K2): rk5_i (the reference's default method, dopri5 with
dense output), rk5_a (controlled dopri5) and rk54_a (controlled Cash-Karp), through every generated
function.  north_star tolerance: final-state agreement within 10 x rtol, accepted-step counts reported.
The GPU and the oracle differ only in pow() of the step-size controller (CUDA libdevice vs glibc, <= 1-2 ulp),
so in practice they agree to ~1e-12 and take identical step counts; the asserted tolerance is the stated one."""
import numpy as np
import pytest

import odeintr_b200
from oracle import build_oracle as bo
from systems import LOGISTIC, LORENZ, LORENZ_PARS, VDP

pytestmark = pytest.mark.gpu


def close(got, ref, rtol, bound=1e-11, scale=None):
    """|got - ref| <= 10 * rtol * (1 + |ref|) elementwise (the stated adaptive-stepper tolerance), and -- beside it -- a
    regression bound: the GPU and the oracle run the same operation sequence and differ only through pow() of the
    controller, so the observed deviations are 1e-15 .. 4e-13; `bound` (in the same norm, plus the same figure relative
    to |ref| alone with a 1e-14 floor for components that pass through zero) catches a kernel that drifts to "merely
    within tolerance".  bound=None: tolerance-level agreement only (Bulirsch-Stoer order selection, fast modes)."""
    import os

    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref) / (1.0 + np.abs(ref))
    worst = float(np.max(err))
    assert worst <= 10 * rtol, worst
    rel = float(np.max(np.abs(got - ref) / (np.abs(ref) + 1e-3 * np.max(np.abs(ref)) + 1e-300)))
    if os.environ.get("ODEINTR_TEST_REPORT"):
        import inspect

        fr = inspect.stack()[1]
        print("close() at %s:%d: err %.3e rel %.3e (10 rtol = %.1e)" % (fr.filename.split("/")[-1], fr.lineno, worst, rel, 10 * rtol))
    if bound is not None:
        assert worst <= bound, worst
        assert rel <= 1e3 * bound, rel
    return worst


def test_default_method_reference_test_facts():
    # tests/testthat/test_odeintr.R:32-46 with the default rk5_i
    env = {}
    code = odeintr_b200.compile_sys("logistic", LOGISTIC, env=env)
    ora = bo.build(LOGISTIC, method="rk5_i")
    df = env["logistic"](0.01, 40)
    assert df.shape == (41, 2) and df.iloc[0, 1] == 0.01 and list(df.columns) == ["Time", "X1"]
    ref = ora.integrate_const([0.01], 0, 40, 1.0)
    assert np.array_equal(df["Time"].to_numpy(), ref["t"])
    close(df["X1"].to_numpy(), ref["rec"][:, 0], 1e-6)
    st = code.system.stats()
    assert st["accepted"][0] == ref["accepted"] == 33 and st["rejected"][0] == ref["rejected"] == 6
    assert st["status"][0] == 0
    with pytest.warns(UserWarning, match="Step size too large -- adjusting"):
        df = env["logistic"](0.01, 40, 100)
    assert df.shape == (4, 2) and df.iloc[0, 1] == 0.01


@pytest.mark.parametrize("method", ["rk5_i", "rk5_a", "rk54_a", "rk78_a", "bs", "bsd"])
def test_config3_vdp_sweep_every_function(method):
    rtol = 1e-8
    env = {}
    code = odeintr_b200.compile_sys("vpol", VDP, "mu", method=method, atol=rtol, rtol=rtol, env=env)
    if method in ("bs", "bsd"):
        # the reference default-constructs Bulirsch-Stoer steppers: tolerances stay 1e-6 whatever was passed, and
        # their order / step-size selection has many thresholds, so a 1-ulp pow() difference occasionally picks
        # another (equally valid) step sequence: agreement is at the tolerance level, not at round-off level
        rtol = 1e-6
    bnd = None if method in ("bs", "bsd") else 1e-9  # regression guard beside the stated tolerance (see close())
    system = code.system
    ora = bo.build(VDP, "mu", False, method, atol=rtol, rtol=rtol)
    m = 257
    mus = 0.5 + 1.5 * np.arange(m) / (m - 1)
    env["vpol_set_params"](mu=mus)
    x0 = np.array([2.0, 0.0])
    X0 = np.repeat(x0[:, None], m, axis=1)

    # name_at: integrate_times with dense output / controlled steps clipped to the sample times
    times = np.arange(0, 20.0001, 0.2)
    rec = env["vpol_at"](x0, times, 0.01)
    e = ora.ensemble(1, X0.copy(), mus[None, :], times=times, dt=0.01, out_rows=len(times), threads=ora.max_threads())
    assert rec.x.shape == (101, 2, m) and np.array_equal(rec.time, times)
    close(rec.x, e["out"], rtol, bound=bnd)
    st = system.stats()
    assert np.all(st["status"] == 0)
    # accepted-step counts are reported and agree with the oracle (a controller decision can flip only when
    # the error lands within an ulp of 1.0)
    same = 0.9 if method in ("bs", "bsd") else 0.98
    assert np.mean(st["accepted"] == e["accepted"]) > same and np.max(np.abs(st["accepted"] - e["accepted"])) <= 3
    assert np.mean(st["rejected"] == e["rejected"]) > same
    close(env["vpol_get_state"](), e["x"], rtol, bound=bnd)

    # name(): integrate_const, observer every step_size
    rec = env["vpol"](x0, 10.0, 0.05)
    e = ora.ensemble(0, X0.copy(), mus[None, :], 0.0, 10.0, 0.05, out_rows=rec.rows, threads=ora.max_threads())
    assert rec.rows == e["rows"]
    close(rec.x, e["out"], rtol, bound=bnd)
    single = ora.integrate_const(x0, 0, 10.0, 0.05, pars=[mus[5]])
    # (dense steppers may or may not deliver the sample at t1 itself, depending on the instance's last real step;
    #  an ensemble record keeps the rows every instance has)
    assert rec.rows in (single["rows"], single["rows"] - 1) and np.array_equal(rec.time, single["t"][:rec.rows])
    df = env["vpol"](x0, 10.0, 0.05)  # the sweep is still set: first instance only?  no: one state x all mu
    assert df.rows == rec.rows

    # name_no_record: integrate_adaptive without observer; state is left for get_state
    fin = env["vpol_no_record"](x0, 7.3, 0.01)
    e = ora.ensemble(2, X0.copy(), mus[None, :], 0.0, 7.3, 0.01, threads=ora.max_threads())
    close(fin, e["x"], rtol, bound=bnd)

    # name_continue_at resumes from the retained state / last recorded time
    env["vpol_at"](x0, np.array([0.0, 1.0, 2.0]), 0.01)
    rec2 = env["vpol_continue_at"](np.array([2.5, 3.0, 4.0]), 0.01)
    r1 = ora.integrate_times(x0, [0.0, 1.0, 2.0], 0.01, pars=[mus[100]])
    adv = ora.integrate_const(r1["x"], 2.0, 2.5, 0.01, pars=[mus[100]], record=False)
    r2 = ora.integrate_times(adv["x"], [2.5, 3.0, 4.0], 0.01, pars=[mus[100]])
    close(rec2.x[:, :, 100], r2["rec"], rtol, bound=bnd)


@pytest.mark.parametrize("method", ["rk5_i", "rk5_a", "rk54_a", "rk78_a", "bs", "bsd"])
def test_adap_ragged_record_single_and_ensemble(method):
    # name_adap: one observer call per accepted step -> every instance has its own Time column
    env = {}
    code = odeintr_b200.compile_sys("lor", LORENZ, LORENZ_PARS, True, method=method, atol=1e-7, rtol=1e-7, env=env)
    ora = bo.build(LORENZ, LORENZ_PARS, True, method, atol=1e-7, rtol=1e-7)
    df = env["lor_adap"](np.ones(3), 3.0, 0.1)
    ref = ora.integrate_adaptive(np.ones(3), 0, 3.0, 0.1, record=True)
    assert len(df) == ref["rows"] and df["Time"].iloc[0] == 0.0 and df["Time"].iloc[-1] == 3.0
    close(df["Time"].to_numpy(), ref["t"], 1e-7, bound=None)
    close(df[["X1", "X2", "X3"]].to_numpy(), ref["rec"], 1e-7 * 100, bound=None)  # Lorenz amplifies 1e-13 differences a little
    rng = np.random.default_rng(5)
    X0 = rng.uniform(-5, 5, (3, 40)) + np.array([[0.0], [0.0], [20.0]])
    rec = env["lor_adap"](X0, 1.0, 0.1)
    for i in (0, 7, 39):
        r = ora.integrate_adaptive(X0[:, i], 0, 1.0, 0.1, record=True)
        n = rec.rows_per_instance[i]
        assert n == r["rows"]
        close(rec.times_per_instance[i, :n], r["t"], 1e-7, bound=None)
        close(rec.x[:n, :, i], r["rec"], 1e-7 * 100, bound=None)


@pytest.mark.parametrize("method,rows", [("rk5_i", 200), ("bsd", 201)])
def test_dense_const_row_count_follows_the_last_real_step(method, rows):
    # integrate_const(0, 10, 0.05) with a dense-output stepper: fl(199 * 0.05) + 0.05 > 10, so Boost's outer loop
    # stops at 9.95; the sample at 10.0 appears only if the inner emission loop runs after the stepper has already
    # reached t1 (Bulirsch-Stoer's big last step does, dopri5 at 1e-8 does not).  A single trajectory must return
    # exactly the reference's rows.
    env = {}
    odeintr_b200.compile_sys("v1", VDP, "mu", method=method, atol=1e-8, rtol=1e-8, env=env)
    ora = bo.build(VDP, "mu", False, method, atol=1e-8, rtol=1e-8)
    env["v1_set_params"](mu=0.535)
    df = env["v1"](np.array([2.0, 0.0]), 10.0, 0.05)
    ref = ora.integrate_const([2.0, 0.0], 0, 10.0, 0.05, pars=[0.535])
    assert len(df) == ref["rows"] == rows
    assert np.array_equal(df["Time"].to_numpy(), ref["t"])
    close(df[["X1", "X2"]].to_numpy(), ref["rec"], 1e-6)


def test_runaway_instances_are_stopped_and_reported():
    # x' = x^2 from x0 = 1 blows up at t = 1.  Boost has no global step limit there (its loop can spin forever on an
    # underflowing dt); the kernels stop such an instance and report it instead of hanging the device.
    env = {}
    code = odeintr_b200.compile_sys("blow", "dxdt = x * x", method="rk5_i", env=env)
    lib = code.system._lib
    assert lib.odeintr_set_max_tries(code.system._h, 5000) == 0
    try:
        env["blow"](np.array([[1.0, 0.5]]), 3.0, 0.1)  # instance 1 blows up at t = 2
        raised = False
    except odeintr_b200.OdeintrError as e:
        raised = "step" in str(e)
    st = code.system.stats()
    assert raised or np.all(st["status"] != 0)
    assert np.all(st["status"] != 0) and np.all(st["accepted"] + st["rejected"] <= 5001)
    # a healthy ensemble next to it is unaffected
    df = env["blow"](0.1, 3.0, 0.1)
    t_last = df["Time"].iloc[-1]  # 2.9: Boost's absolute-epsilon loop drops the sample at fl(30 * 0.1) > 3
    assert abs(df["X1"].iloc[-1] - 0.1 / (1 - 0.1 * t_last)) < 1e-5


@pytest.mark.parametrize("method", ["rk5_i", "rk5_a"])
def test_continue_at_after_adap_single_and_ensemble(method):
    # NAME_adap() leaves a ragged record (one row per accepted step); NAME_continue_at() must continue from
    # its last time like the reference (rec_t.back(), compile_sys_template.cpp:128-139) instead of failing
    rtol = 1e-8
    env = {}
    code = odeintr_b200.compile_sys("vp", VDP, "mu", method=method, atol=rtol, rtol=rtol, env=env)
    ora = bo.build(VDP, "mu", False, method, atol=rtol, rtol=rtol)
    env["vp_set_params"](mu=1.3)
    x0 = np.array([2.0, 0.0])
    df = env["vp_adap"](x0, 3.0, 0.01)
    assert df["Time"].iloc[-1] == 3.0
    more = np.array([3.0, 3.5, 4.25])
    df2 = env["vp_continue_at"](more, 0.01)
    a = ora.integrate_adaptive(x0, 0.0, 3.0, 0.01, pars=[1.3])
    ref = ora.integrate_times(a["x"], more, 0.01, pars=[1.3])
    assert np.array_equal(df2["Time"].to_numpy(), more)
    err = close(df2[["X1", "X2"]].to_numpy(), ref["rec"], rtol)
    assert err <= 1e-11, err
    # ensemble: every instance ends at start + duration, so the sweep continues from there as well
    mus = np.linspace(0.5, 2.0, 37)
    env["vp_set_params"](mu=mus)
    rec = env["vp_adap"](x0, 2.0, 0.01)
    assert len(set(rec.rows_per_instance.tolist())) > 1  # genuinely ragged
    rec2 = env["vp_continue_at"](np.array([2.0, 2.5]), 0.01)
    X = np.repeat(x0[:, None], mus.size, axis=1)
    e1 = ora.ensemble(2, X, mus[None, :], 0.0, 2.0, 0.01, threads=ora.max_threads())
    e2 = ora.ensemble(1, e1["x"], mus[None, :], times=np.array([2.0, 2.5]), dt=0.01, out_rows=2, threads=ora.max_threads())
    assert close(rec2.x, e2["out"], rtol) <= 1e-11
