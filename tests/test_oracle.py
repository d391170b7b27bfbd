"""The oracle (CPU restatement of Boost odeint) against everything that pins it: SURVEY.md Appendix B
golden bit patterns, the reference's own test facts (tests/testthat/test_odeintr.R:25-54: 41 rows, 4 rows,
first row == initial state), analytic solutions and convergence orders."""
import json
import math
import pathlib
import struct

import numpy as np
import pytest

from oracle import build_oracle as bo
from systems import LOGISTIC, LORENZ, LORENZ_PARS, ROBERTSON, ROBERTSON_PARS, STIFF, VDP

GOLDEN = json.loads((pathlib.Path(__file__).parent / "golden" / "lorenz_rk4_golden.json").read_text())


def hexd(v):
    return struct.pack(">d", float(v)).hex()


def test_lorenz_rk4_golden_bits():
    const = bo.build(LORENZ, LORENZ_PARS, True, "rk4")
    dyn = bo.build(LORENZ, LORENZ_PARS, False, "rk4")
    runs = {"const_text": const.integrate_const([1, 1, 1], 0, 100, 0.001),
            "exact_8_3": dyn.integrate_const([1, 1, 1], 0, 100, 0.001, pars=[10, 28, 8 / 3])}
    for case in GOLDEN["cases"]:
        r = runs[case["pars"]]
        assert r["rows"] == 100001
        for n, bits in case["steps"].items():
            assert [hexd(v) for v in r["rec"][int(n)]] == bits, (case["pars"], n)


def test_const_step_counts_follow_boost_loop():
    # absolute-epsilon comparison drops the last step for some (t1, dt)
    o = bo.build(LOGISTIC, method="rk4")
    for t1, dt, rows in [(100, 0.001, 100001), (40, 1, 41), (40, 40 / 3, 4), (15, 0.01, 1501), (100, 0.01, 10000),
                         (7, 0.001, 7000), (20, 0.01, 2000)]:
        r = o.integrate_const([0.01], 0, t1, dt)
        assert r["rows"] == rows, (t1, dt)


def logistic(t, x0=0.01):
    return x0 * math.exp(t) / (1 + x0 * (math.exp(t) - 1))


def test_reference_test_facts_default_method():
    # test_odeintr.R:32-46: compile_sys("logistic", "dxdt = x * (1 - x)"); logistic(0.01, 40) -> 41 x 2
    o = bo.build(LOGISTIC, method="rk5_i")
    r = o.integrate_const([0.01], 0, 40, 1.0)
    assert r["rows"] == 41 and r["rec"][0, 0] == 0.01 and r["t"][0] == 0.0
    assert r["accepted"] == 33 and r["rejected"] == 6  # SURVEY.md A.3 re-simulation
    for t, x in zip(r["t"], r["rec"][:, 0]):
        assert abs(x - logistic(t)) < 2e-5
    # large step: the host guard turns step 100 into 40/3 -> 4 rows
    r = o.integrate_const([0.01], 0, 40, 40 / 3)
    assert r["rows"] == 4 and r["rec"][0, 0] == 0.01


def test_reference_test_facts_implicit():
    # test_odeintr.R:48-54 through rosenbrock4
    o = bo.build("dxdt[0] = x[0] * (1 - x[0])", method="rosenbrock4")
    r = o.integrate_const([0.01], 0, 40, 1.0)
    assert r["rows"] == 41 and r["rec"][0, 0] == 0.01
    for t, x in zip(r["t"], r["rec"][:, 0]):
        assert abs(x - logistic(t)) < 5e-5
    r = o.integrate_const([0.01], 0, 40, 40 / 3)
    assert r["rows"] == 4


@pytest.mark.parametrize("method,order", [("euler", 1), ("rk4", 4), ("rk54", 5), ("rk5", 5), ("rk78", 8)])
def test_convergence_order(method, order):
    o = bo.build("dxdt[0] = x[1]; dxdt[1] = -x[0]", method=method)
    errs = []
    for n in ((4, 8) if order == 8 else (50, 100)):
        r = o.integrate_const([1.0, 0.0], 0, 1.0, 1.0 / n, record=False)
        errs.append(math.hypot(r["x"][0] - math.cos(1.0), r["x"][1] + math.sin(1.0)))
    assert abs(math.log2(errs[0] / errs[1]) - order) < 0.35


def test_stiff_linear_system_analytic():
    # README.md:173-176: eigenvalues -1 and -100; x0 = (2, 1) -> x1(t) = e^-t-ish combination
    o = bo.build(STIFF, method="rosenbrock4")
    r = o.integrate_const([2.0, 1.0], 0, 5, 0.001)
    assert r["rows"] == 5001
    # exact solution via eigen-decomposition
    A = np.array([[-101.0, -100.0], [1.0, 0.0]])
    w, V = np.linalg.eig(A)
    c = np.linalg.solve(V, np.array([2.0, 1.0]))
    for k in (0, 1, 10, 1000, 5000):
        t = r["t"][k]
        exact = (V * np.exp(w * t)) @ c
        assert np.allclose(r["rec"][k], exact.real, rtol=0, atol=2e-5)


def test_robertson_conservation_and_times():
    o = bo.build(ROBERTSON, ROBERTSON_PARS, True, "rosenbrock4")
    times = 10 ** np.linspace(-5, 5, 41)
    r = o.integrate_times([1.0, 0.0, 0.0], times, 1.0)
    assert r["rows"] == 41
    assert np.allclose(r["rec"].sum(axis=1), 1.0, atol=1e-4)
    assert np.all(r["rec"] >= -1e-6)
    # SciPy cross-check
    from scipy.integrate import solve_ivp

    def f(t, y):
        a, b, g = 0.04, 1e4, 3e7
        return [-a * y[0] + b * y[1] * y[2], a * y[0] - b * y[1] * y[2] - g * y[1] ** 2, g * y[1] ** 2]

    s = solve_ivp(f, (0, times[-1]), [1, 0, 0], method="Radau", t_eval=times[times > 0], rtol=1e-10, atol=1e-12)
    assert np.allclose(r["rec"][:, 0], s.y[0], rtol=0, atol=5e-5)
    assert np.allclose(r["rec"][:, 2], s.y[2], rtol=0, atol=5e-5)


def test_vdp_dense_times_against_scipy():
    from scipy.integrate import solve_ivp

    o = bo.build(VDP, "mu", False, "rk5_i", atol=1e-8, rtol=1e-8)
    times = np.arange(0, 20.0001, 0.2)
    r = o.integrate_times([2.0, 0.0], times, 0.01, pars=[1.5])
    assert r["rows"] == len(times)
    s = solve_ivp(lambda t, y: [y[1], 1.5 * (1 - y[0] ** 2) * y[1] - y[0]], (0, 20), [2.0, 0.0], method="DOP853",
                  t_eval=times, rtol=1e-13, atol=1e-13)
    assert np.max(np.abs(r["rec"] - s.y.T)) < 1e-5


def test_integrate_adaptive_variants_reach_t1():
    for method in ("rk4", "rk5_a", "rk5_i", "rk54_a", "rk78_a", "bs", "bsd"):
        o = bo.build(LOGISTIC, method=method)
        r = o.integrate_adaptive([0.01], 0, 7.3, 0.25, record=True)
        assert r["t"][-1] == 7.3
        assert abs(r["x"][0] - logistic(7.3)) < 1e-4, method


def test_ensemble_matches_single_runs():
    o = bo.build(LORENZ, LORENZ_PARS, False, "rk4")
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (3, 7))
    P = np.stack([np.full(7, 10.0), np.linspace(20, 30, 7), np.full(7, 8 / 3)])
    e = o.ensemble(0, X.copy(), P, 0, 0.5, 0.01, obs_stride=10, out_rows=6, threads=2)
    assert e["rows"] == 6
    for i in range(7):
        r = o.integrate_const(X[:, i], 0, 0.5, 0.01, pars=P[:, i])
        assert np.array_equal(e["out"][:, :, i], r["rec"][::10])
        assert np.array_equal(e["x"][:, i], r["x"])


def test_bulirsch_stoer_dense_output_is_exact_for_polynomial_solutions():
    # bsd interpolates with the Taylor polynomial about the interval centre built from extrapolated finite
    # differences: for x(t) = t^3 + 2t every sample between the (5) real steps must be exact to round-off, which
    # pins the indexing and the (dt/2)^k / k! scaling of the restated prepare_dense_output / calc_state
    o = bo.build("dxdt[0] = 3.0 * x[1] * x[1] + 2.0; dxdt[1] = 1.0", method="bsd")
    r = o.integrate_const([0.0, 0.0], 0, 10, 0.1)
    t = r["t"]
    assert r["rows"] == 101 and r["accepted"] <= 8
    assert np.max(np.abs(r["rec"][:, 0] - (t ** 3 + 2 * t))) < 1e-10 and np.max(np.abs(r["rec"][:, 1] - t)) < 1e-12
    # README.md:125-132: vanderpol(rep(1e-4, 2), 100, 0.01) with bsd -> 10 001 observer calls
    o = bo.build(VDP, "mu", False, "bsd")
    r = o.integrate_const([1e-4, 1e-4], 0, 100, 0.01, pars=[1.0])
    assert r["rows"] == 10001


def test_dopri5_and_rk4_single_steps_against_independent_implementations():
    # an implementation the restatement does not share a line with: SciPy's Dormand-Prince tableau (RK45.A/B/C) driven
    # by scipy.integrate._ivp.rk.rk_step, and the textbook RK4 formula in numpy.  Different summation orders, so the
    # comparison is to a few ulps, over several steps of a nonlinear system
    from scipy.integrate._ivp import rk as sp_rk

    def f(t, y):
        return np.array([y[1], 1.5 * (1 - y[0] * y[0]) * y[1] - y[0] + 0.1 * t])

    src = "dxdt[0] = x[1]; dxdt[1] = 1.5 * (1 - x[0] * x[0]) * x[1] - x[0] + 0.1 * t;"
    h = 0.05
    o5 = bo.build(src, method="rk5")
    o4 = bo.build(src, method="rk4")
    y5 = y4 = np.array([1.2, -0.4])
    t = 0.0
    for step in range(8):
        K = np.empty((7, 2))
        y5_new, _ = sp_rk.rk_step(f, t, y5, f(t, y5), h, sp_rk.RK45.A, sp_rk.RK45.B, sp_rk.RK45.C, K)
        k1 = f(t, y4); k2 = f(t + h / 2, y4 + h / 2 * k1); k3 = f(t + h / 2, y4 + h / 2 * k2); k4 = f(t + h, y4 + h * k3)
        y4_new = y4 + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        r5 = o5.integrate_const(y5, t, t + 1.5 * h, h, record=False)  # exactly one step of h
        r4 = o4.integrate_const(y4, t, t + 1.5 * h, h, record=False)
        assert np.max(np.abs(r5["x"] - y5_new)) < 4e-15, (step, r5["x"], y5_new)
        assert np.max(np.abs(r4["x"] - y4_new)) < 4e-15, (step, r4["x"], y4_new)
        y5, y4, t = r5["x"], r4["x"], t + h


def test_dopri5_error_estimate_and_controller_against_scipy_tableau():
    # integrate_adaptive with rk5_a: the accepted time grid depends on the embedded error estimate and on Boost's
    # default_error_checker / default_step_adjuster.  Rebuilt here from SciPy's RK45 tableau and error weights (RK45.E)
    # plus the controller formulas of SURVEY.md A.2; the restated oracle must walk the same grid
    from scipy.integrate._ivp import rk as sp_rk

    def f(t, y):
        return np.array([y[1], 2.0 * (1 - y[0] * y[0]) * y[1] - y[0]])

    src = "dxdt[0] = x[1]; dxdt[1] = 2.0 * (1 - x[0] * x[0]) * x[1] - x[0];"
    atol = rtol = 1e-6
    o = bo.build(src, method="rk5_a", atol=atol, rtol=rtol)
    t1, dt = 6.0, 0.1
    ref = o.integrate_adaptive([2.0, 0.0], 0.0, t1, dt, record=True)
    t, y, grid, rejected = 0.0, np.array([2.0, 0.0]), [0.0], 0
    while t < t1 - 1e-15 * max(abs(t), abs(t1)):
        if t + dt > t1:
            dt = t1 - t
        while True:
            K = np.empty((7, 2))
            d = f(t, y)
            y_new, _ = sp_rk.rk_step(f, t, y, d, dt, sp_rk.RK45.A, sp_rk.RK45.B, sp_rk.RK45.C, K)
            xerr = dt * (K.T @ sp_rk.RK45.E)
            err = np.max(np.abs(xerr) / (atol + rtol * (np.abs(y) + abs(dt) * np.abs(d))))
            if err > 1.0:
                dt *= max(0.9 * err ** (-1.0 / 3.0), 0.2)
                rejected += 1
                continue
            t += dt
            if err < 0.5:
                dt *= 0.9 * max(5.0 ** -5, err) ** (-1.0 / 5.0)
            y = y_new
            break
        grid.append(t)
    assert len(grid) == ref["rows"] and rejected == ref["rejected"], (len(grid), ref["rows"], rejected, ref["rejected"])
    assert np.max(np.abs(np.array(grid) - ref["t"])) < 1e-9
    assert np.max(np.abs(y - ref["rec"][-1])) < 1e-9


# ---- Rosenbrock4: pins for the ~45 Ross constants shared by the oracle and the K3 kernel ----------
def _ross_constants(path, begin, end):
    """name -> value for every `name = number` between two markers of a source file."""
    import re

    text = pathlib.Path(path).read_text()
    block = text[text.index(begin):]
    block = block[:block.index(end)]
    return {k: float(v) for k, v in re.findall(r"\b([acd]\d{1,2}|gamma)\s*=\s*(-?[0-9.]+(?:e[+-]?[0-9]+)?)", block)}


ROOT = pathlib.Path(__file__).resolve().parent.parent
ROSS_SOURCES = {
    "oracle": (ROOT / "oracle" / "odeint_restated.hpp", "struct rosenbrock4 {", "state_type dxdt, dfdt"),
    "kernel": (ROOT / "odeintr_b200" / "csrc" / "kernels" / "rosenbrock4.cuh", "struct ross_coef {", "};"),
}


@pytest.mark.parametrize("which", ["oracle", "kernel"])
def test_rosenbrock4_constants_satisfy_the_order_conditions(which):
    # Boost's rosenbrock4 works with the transformed increments g_i = sum_j gamma_ij k_j (no matrix-vector products):
    # A = alpha Gamma^-1, C = diag(1/gamma) - Gamma^-1, m = b Gamma^-1.  Undo the transformation and check the classical
    # order conditions of Rosenbrock methods (Hairer & Wanner, Solving ODEs II, IV.7, Table 7.1): order 4 for the
    # solution weights, order 3 for the embedded ones, and the consistency of the non-autonomous terms.  A typo in any
    # of the a / c constants breaks at least one of them by far more than the 1e-12 the published digits allow.
    k = _ross_constants(*ROSS_SOURCES[which])
    g = k["gamma"]
    s = 6
    A = np.zeros((s, s))
    Cm = np.zeros((s, s))
    for i in range(2, 6):
        for j in range(1, i):
            A[i - 1, j - 1] = k["a%d%d" % (i, j)]
    A[5, :] = [k["a51"], k["a52"], k["a53"], k["a54"], 1.0, 0.0]  # stage 6 is evaluated at xtmp + g5
    for i in range(2, 7):
        for j in range(1, i):
            Cm[i - 1, j - 1] = k["c%d%d" % (i, j)]
    m = np.array([k["a51"], k["a52"], k["a53"], k["a54"], 1.0, 1.0])   # x_new = x + sum m_i g_i
    mh = np.array([k["a51"], k["a52"], k["a53"], k["a54"], 1.0, 0.0])  # embedded: without g6 (= the error estimate)
    Gamma = np.linalg.inv(np.eye(s) / g - Cm)
    alpha = A @ Gamma
    b, bh = m @ Gamma, mh @ Gamma
    assert np.allclose(np.diag(Gamma), g, atol=1e-14) and np.allclose(np.triu(alpha), 0, atol=1e-13)
    beta = alpha + Gamma - g * np.eye(s)  # beta_ij = alpha_ij + gamma_ij, j < i
    ai, bi = alpha.sum(axis=1), beta.sum(axis=1)
    tol = 2e-12
    for w, order in ((b, 4), (bh, 3)):
        assert abs(w.sum() - 1) < tol
        assert abs(w @ bi - (0.5 - g)) < tol
        assert abs(w @ ai ** 2 - 1 / 3) < tol
        assert abs(w @ (beta @ bi) - (1 / 6 - g + g * g)) < tol
        if order == 4:
            assert abs(w @ ai ** 3 - 0.25) < tol
            assert abs(w @ (ai * (alpha @ bi)) - (1 / 8 - g / 3)) < tol
            assert abs(w @ (beta @ ai ** 2) - (1 / 12 - g / 3)) < tol
            assert abs(w @ (beta @ (beta @ bi)) - (1 / 24 - g / 2 + 1.5 * g * g - g ** 3)) < tol
    # the embedded weights are NOT of order 4 (otherwise the error estimate would be useless)
    assert abs(bh @ ai ** 3 - 0.25) > 1e-3 or abs(bh @ (beta @ (beta @ bi)) - (1 / 24 - g / 2 + 1.5 * g * g - g ** 3)) > 1e-3
    # non-autonomous terms: c_i = sum_j alpha_ij, d_i = gamma + sum_j gamma_ij (rows 5 and 6 have c = 1, d = 0)
    assert np.allclose(ai[1:4], [k["c2"], k["c3"], k["c4"]], atol=1e-12) and np.allclose(ai[4:], 1.0, atol=1e-12)
    assert np.allclose(Gamma.sum(axis=1)[:4], [k["d1"], k["d2"], k["d3"], k["d4"]], atol=1e-12)
    assert np.allclose(Gamma.sum(axis=1)[4:], 0.0, atol=1e-12)


def test_rosenbrock4_constants_identical_in_oracle_and_kernel():
    a, b = (_ross_constants(*ROSS_SOURCES[w]) for w in ("oracle", "kernel"))
    assert len(a) == 43 and a == b  # gamma, d1-d4, c2-c4, 15 c_ij, 10 a_ij, 10 dense-output d_ij


def test_rosenbrock4_fixed_step_orders_through_the_oracle_arithmetic():
    # one raw do_step of the oracle at h and h/2 on a nonlinear stiff-ish system with a known solution:
    # local error of the 4th-order result ~ h^5 (ratio 32), of the embedded 3rd-order result ~ h^4 (ratio 16),
    # of the dense output at mid-step ~ h^4; global error over a fixed-step integration ~ h^4
    src = "dxdt[0] = -2.0 * x[0] + x[1] * x[1]; dxdt[1] = -x[1];"   # x1 = e^-t, x0 = (1 + t) e^-2t for x(0) = (1, 1)
    o = bo.build(src, method="rosenbrock4")

    def exact(t):
        return np.array([(1 + t) * math.exp(-2 * t), math.exp(-t)])

    e4, e3, ed = [], [], []
    for h in (0.1, 0.05):
        xn, xe, xd = o.rosenbrock4_step(exact(0.3), 0.3, h, 0.5)
        e4.append(np.linalg.norm(xn - exact(0.3 + h)))
        e3.append(np.linalg.norm((xn - xe) - exact(0.3 + h)))
        ed.append(np.linalg.norm(xd - exact(0.3 + 0.5 * h)))
    assert abs(math.log2(e4[0] / e4[1]) - 5) < 0.4, e4
    assert abs(math.log2(e3[0] / e3[1]) - 4) < 0.4, e3
    assert abs(math.log2(ed[0] / ed[1]) - 4) < 0.5, ed
    # dense output is continuous: theta = 0 / 1 give the step's end points
    x0 = exact(0.3)
    xn, _, d0 = o.rosenbrock4_step(x0, 0.3, 0.1, 0.0)
    _, _, d1 = o.rosenbrock4_step(x0, 0.3, 0.1, 1.0)
    assert np.allclose(d0, x0, atol=1e-15) and np.allclose(d1, xn, atol=1e-15)
    errs = []
    for n in (20, 40):
        x = exact(0.0)
        for s in range(n):
            x, _, _ = o.rosenbrock4_step(x, s / n, 1.0 / n)
        errs.append(np.linalg.norm(x - exact(1.0)))
    assert abs(math.log2(errs[0] / errs[1]) - 4) < 0.35, errs
