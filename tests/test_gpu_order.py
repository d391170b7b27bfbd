"""K6: instance order of adaptive ensembles (odeintr_set_instance_order).  Sorting the instances by the pilot's cost
proxy (or by a caller's permutation) changes which trajectories share a warp and nothing else: per-instance results and
step counts must be IDENTICAL, bit for bit, to the run in index order (north_star: early-finishing / heterogeneous
adaptive instances; the reference's serial loop over instances is README.md:157-163)."""
import numpy as np
import pytest

import odeintr_b200
from oracle import build_oracle as bo
from systems import LORENZ, LORENZ_PARS, ROBERTSON, ROBERTSON_PARS, VDP

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("method", ["rk5_i", "rk5_a", "rk78_a", "bsd"])
def test_shuffled_vdp_sweep_identical_results_in_every_order(method):
    rtol = 1e-8
    env = {}
    code = odeintr_b200.compile_sys("vp", VDP, "mu", method=method, atol=rtol, rtol=rtol, env=env)
    s = code.system
    m = 4099  # ragged last warp
    mus = np.random.default_rng(5).uniform(0.1, 10.0, m)
    env["vp_set_params"](mu=mus)
    times = np.arange(0, 8.0001, 0.2)
    x0 = np.array([2.0, 0.0])
    runs = {}
    for order in ("index", "cost", np.random.default_rng(6).permutation(m)):
        s.set_instance_order(order)
        rec = env["vp_at"](x0, times, 0.01)
        st = s.stats()
        fin = env["vp_get_state"]()
        key = order if isinstance(order, str) else "perm"
        runs[key] = (rec.x.copy(), st["accepted"].copy(), st["rejected"].copy(), fin.copy())
    perm = None
    s.set_instance_order("cost")
    env["vp_at"](x0, times, 0.01)
    perm = s.get_instance_order()
    assert np.array_equal(np.sort(perm), np.arange(m))  # a permutation of the instances ...
    assert not np.array_equal(perm, np.arange(m))       # ... that is not the identity for a shuffled sweep
    # ... and that orders them by stiffness: mu along the sorted order is (nearly) monotone
    # (dopri5's controller has settled after the pilot's 12 attempts; the high-order / extrapolation steppers take
    #  larger first steps and order the sweep less sharply, which costs speed, never correctness)
    rank_corr = np.corrcoef(np.argsort(np.argsort(mus[perm])), np.arange(m))[0, 1]
    if method in ("rk5_i", "rk5_a"):
        assert abs(rank_corr) > 0.9, rank_corr
    for key in ("cost", "perm"):
        for a, b in zip(runs["index"], runs[key]):
            assert np.array_equal(bits(a) if a.dtype == np.float64 else a, bits(b) if b.dtype == np.float64 else b), key
    # and the run in index order is the oracle's, as everywhere else
    ora = bo.build(VDP, "mu", False, method, atol=1e-6 if method == "bsd" else rtol, rtol=1e-6 if method == "bsd" else rtol)
    idx = np.arange(0, m, 97)
    e = ora.ensemble(1, np.repeat(x0[:, None], idx.size, axis=1), mus[None, idx], times=times, dt=0.01, out_rows=times.size,
                     threads=ora.max_threads())
    err = np.max(np.abs(runs["cost"][0][:, :, idx] - e["out"]) / (1 + np.abs(e["out"])))
    assert err <= 10 * (1e-6 if method == "bsd" else rtol), err


def test_random_lorenz_and_ragged_adap_record_identical_in_cost_order():
    env = {}
    code = odeintr_b200.compile_sys("lor", LORENZ, LORENZ_PARS, True, method="rk5_i", atol=1e-8, rtol=1e-8, env=env)
    s = code.system
    m = 2500
    rng = np.random.default_rng(7)
    X0 = np.stack([rng.uniform(-20, 20, m), rng.uniform(-25, 25, m), rng.uniform(5, 45, m)])
    out = {}
    for order in ("index", "cost"):
        s.set_instance_order(order)
        rec = env["lor_at"](X0, np.arange(0, 3.0001, 0.1), 0.01)
        st = s.stats()
        rag = env["lor_adap"](X0[:, :60], 1.0, 0.01)  # ragged record: the dry run and the recording run share the order
        out[order] = (rec.x.copy(), st["accepted"].copy(), st["rejected"].copy(), rag.x.copy(), rag.rows_per_instance.copy())
    for a, b in zip(out["index"], out["cost"]):
        assert np.array_equal(bits(a) if a.dtype == np.float64 else a, bits(b) if b.dtype == np.float64 else b)


def test_cost_order_for_the_implicit_stepper_and_error_cases():
    env = {}
    code = odeintr_b200.compile_implicit("rob", ROBERTSON, ROBERTSON_PARS, env=env)
    s = code.system
    m = 1500
    alphas = np.random.default_rng(8).uniform(0.02, 0.08, m)
    env["rob_set_params"](alpha=alphas)
    times = 10 ** np.linspace(-5, 3, 17)
    res = {}
    for order in ("index", "cost"):
        s.set_instance_order(order)
        rec = env["rob_at"](np.array([1.0, 0.0, 0.0]), times)
        st = s.stats()
        res[order] = (rec.x.copy(), st["accepted"].copy(), st["rejected"].copy())
    for a, b in zip(res["index"], res["cost"]):
        assert np.array_equal(bits(a) if a.dtype == np.float64 else a, bits(b) if b.dtype == np.float64 else b)
    with pytest.raises(odeintr_b200.OdeintrError, match="Invalid instance order"):
        s.set_instance_order("fastest")
    with pytest.raises(odeintr_b200.OdeintrError, match="not a permutation"):
        s.set_instance_order(np.array([0, 0, 1]))
    # a permutation made for another ensemble size is ignored (index order), not misapplied
    s.set_instance_order(np.arange(10)[::-1].copy())
    rec = env["rob_at"](np.array([1.0, 0.0, 0.0]), times)
    assert np.array_equal(bits(rec.x), bits(res["index"][0]))
