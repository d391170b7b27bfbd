"""GPU parity for large coupled systems (K4, compile_sys_openmp, BASELINE.json configs[4]): fused
Runge-Kutta stage kernels against the oracle running the same loops on the CPU.  Fixed-step arithmetic in
Boost's order without FMA -> bit-exact."""
import numpy as np
import pytest

import odeintr_b200
from oracle import build_oracle as bo
from systems import BISTABLE_1D, BISTABLE_2D, BISTABLE_DOC, BISTABLE_PARS, CHAIN, CHAIN_GLOBALS, CHAIN_PARS

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def cols(df):
    return df.iloc[:, 1:].to_numpy()


@pytest.mark.parametrize("method,halo", [("rk4", None), ("rk4", 0), ("euler", None)])
def test_bistable_chain_bit_exact(method, halo):
    # halo=None: the drop-in call; for rk4 the stencil radius is measured and the temporally blocked kernels run.
    # halo=0: the stage-per-launch kernels.
    n = 4099  # not a multiple of anything: ragged last block / tile, `% N` wrap at both ends
    env = {}
    code = odeintr_b200.compile_sys_openmp("bistable", BISTABLE_1D, BISTABLE_PARS, True, method=method, sys_dim=n,
                                           env=env, halo=halo)
    assert code.system.get_halo() == (1 if (method, halo) == ("rk4", None) else -1)
    ora = bo.build(BISTABLE_1D, BISTABLE_PARS, True, method, sys_dim=n)
    rng = np.random.default_rng(11)
    x0 = rng.integers(0, 2, n).astype(np.float64)  # inic = rbinom(M * M, 1, 1/2) in the reference's example
    df = env["bistable"](x0, 5.0, 0.1)
    ref = ora.integrate_const(x0, 0, 5.0, 0.1)
    assert df.shape == (ref["rows"], n + 1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"])) and np.array_equal(df["Time"].to_numpy(), ref["t"])
    # bistable_at(inic, at) of the example
    at = np.array([1.0, 2.5, 10.0 ** 0.5, 7.0])
    df = env["bistable_at"](x0, at, 0.25)
    adv = ora.integrate_const(x0, 0.0, at[0], 0.25, record=False)
    ref = ora.integrate_times(adv["x"], at, 0.25)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    fin = env["bistable_no_record"](x0, 3.3, 0.25)
    ref = ora.integrate_adaptive(x0, 0, 3.3, 0.25)
    assert np.array_equal(bits(fin), bits(ref["x"]))
    assert np.array_equal(bits(env["bistable_get_state"]()), bits(ref["x"]))


def test_two_loops_and_laplace2d_helper():
    m = 48
    env = {}
    odeintr_b200.compile_sys_openmp("b2d", BISTABLE_2D, BISTABLE_PARS, True, method="rk4", sys_dim=m * m, env=env)
    ora = bo.build(BISTABLE_2D, BISTABLE_PARS, True, "rk4", sys_dim=m * m)
    rng = np.random.default_rng(12)
    x0 = rng.integers(0, 2, m * m).astype(np.float64)
    df = env["b2d"](x0, 4.0, 0.2)
    ref = ora.integrate_const(x0, 0, 4.0, 0.2)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))


@pytest.mark.parametrize("halo", [None, 0])
def test_oscillator_chain_two_fields_runtime_parameters(halo):
    n = 2 * 3001
    env = {}
    code = odeintr_b200.compile_sys_openmp("chain", CHAIN, CHAIN_PARS, False, method="rk4", sys_dim=n,
                                           globals=CHAIN_GLOBALS, env=env, halo=halo)
    ora = bo.build(CHAIN, CHAIN_PARS, False, "rk4", sys_dim=n, globals=CHAIN_GLOBALS)
    L = n // 2
    x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(L) / L), np.zeros(L)])
    env["chain_set_params"](beta=0.7, K=0.45)
    df = env["chain"](x0, 2.0, 0.01)
    assert code.system.get_halo() == (1 if halo is None else -1)
    ref = ora.integrate_const(x0, 0, 2.0, 0.01, pars=[0.7, 0.45])
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    # the symplectic-ish energy drifts only slightly over 200 rk4 steps: a size-independent sanity property
    def energy(s):
        q, p = s[:L], s[L:]
        return np.sum(0.5 * p * p + 0.5 * q * q + 0.7 * q ** 4 / 4 + 0.45 * 0.5 * (np.roll(q, -1) - q) ** 2)
    e0, e1 = energy(x0), energy(cols(df)[-1])
    assert abs(e1 - e0) / e0 < 1e-8


def test_large_system_errors():
    with pytest.raises(odeintr_b200.OdeintrError, match="sys_dim"):
        odeintr_b200.compile_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, method="rk4")
    with pytest.raises(odeintr_b200.OdeintrError, match="Invalid integration method"):
        odeintr_b200.compile_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, method="rk4_a", sys_dim=64)


@pytest.mark.parametrize("case", ["bistable", "chain"])
def test_sharded_stepping_single_rank_matches_unsharded(case):
    # odeintr_lattice_shard_step with world = 1: ghost cells wrap onto the same slab; the shrinking-window
    # schedule must give the bits of the plain single-GPU path (and hence of the oracle)
    import torch

    from odeintr_b200.distributed import ShardedLattice

    if case == "bistable":
        n_cells, fields = 5003, 1
        code = odeintr_b200.compile_sys_openmp("bs", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n_cells)
        code2 = odeintr_b200.compile_sys_openmp("bs2", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n_cells)
        rng = np.random.default_rng(31)
        x0 = rng.integers(0, 2, n_cells).astype(np.float64)
    else:
        n_cells, fields = 4001, 2
        code = odeintr_b200.compile_sys_openmp("ch", CHAIN, CHAIN_PARS, True, method="rk4", sys_dim=2 * n_cells,
                                               globals=CHAIN_GLOBALS)
        code2 = odeintr_b200.compile_sys_openmp("ch2", CHAIN, CHAIN_PARS, True, method="rk4", sys_dim=2 * n_cells,
                                                globals=CHAIN_GLOBALS)
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(n_cells) / n_cells), np.zeros(n_cells)])
    ref = code.system.no_record(x0, 2.0, 0.1)  # 20 rk4 steps on the unsharded path
    lat = ShardedLattice(code2.system, n_cells, fields=fields, halo=1, rank=0, world=1, native=(case == "chain"))
    lat.set_state_from(lambda f, cells: x0[f * n_cells + cells])
    lat.step(12, 0.1, t0=0.0)
    lat.step(8, 0.1)  # continuing keeps the integrate_const time rule t0 + step * dt
    torch.cuda.synchronize()
    got = lat.gather_state().cpu().numpy().reshape(-1)
    assert np.array_equal(bits(got), bits(ref))


@pytest.mark.parametrize("world,case", [(2, "bistable"), (3, "bistable"), (2, "chain"), (4, "chain"), (2, "lattice2d"),
                                        (3, "lattice2d")])
def test_sharded_stepping_emulated_ranks_on_one_gpu(world, case):
    # several slabs of one lattice inside one process: the halo exchange is done by hand between the slabs'
    # tensors, the kernels are the multi-GPU ones (window below 0 / above n_total = periodic images; interior
    # chunks on shifted plain pointers, one shift per field)
    import torch

    from odeintr_b200.distributed import ShardedLattice

    if case == "bistable":
        n_cells, fields = 3001, 1
        rng = np.random.default_rng(41)
        x0 = rng.integers(0, 2, n_cells).astype(np.float64)
        build = lambda name: odeintr_b200.compile_sys_openmp(name, BISTABLE_1D, BISTABLE_PARS, True, method="rk4",
                                                             sys_dim=n_cells).system
    elif case == "lattice2d":
        # an M x M lattice (laplace2D) in row-block slabs: in the flat row-major numbering the vertical neighbours are
        # M cells away, so the slabs are 1-D slabs with a stencil radius of M cells (whole ghost rows); slab boundaries
        # need not fall on row boundaries (96 * 96 cells over 2 or 3 ranks)
        side = 96
        n_cells, fields, halo = side * side, 1, side
        x0 = np.random.default_rng(47).uniform(0, 1, n_cells)
        build = lambda name: odeintr_b200.compile_sys_openmp(name, BISTABLE_2D, BISTABLE_PARS, True, method="rk4",
                                                             sys_dim=n_cells).system
    else:
        n_cells, fields = 2 * 4099, 2  # cells per field; enough for several chunks per slab
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(n_cells) / n_cells), np.zeros(n_cells)])
        build = lambda name: odeintr_b200.compile_sys_openmp(name, CHAIN, CHAIN_PARS, True, method="rk4",
                                                             sys_dim=2 * n_cells, globals=CHAIN_GLOBALS).system
    ref = build("ref").no_record(x0, 1.5, 0.1)  # 15 steps
    lats = []
    for r in range(world):
        lat = ShardedLattice(build("r%d" % r), n_cells, fields=fields, rank=r, world=world, native=False,
                             **({"halo": halo} if case == "lattice2d" else {}))
        lat.set_state_from(lambda f, cells: x0[f * n_cells + cells])
        lats.append(lat)
    g, pad = lats[0].ghost, lats[0].pad
    for step in range(15):
        for lat in lats:
            lat.stream.synchronize()
        for r, lat in enumerate(lats):
            left, right = lats[(r - 1) % world], lats[(r + 1) % world]
            lat.x[:, pad - g:pad] = left.x[:, pad + left.n_local - g:pad + left.n_local]
            lat.x[:, pad + lat.n_local:pad + lat.n_local + g] = right.x[:, pad:pad + g]
        torch.cuda.synchronize()
        for lat in lats:
            lat.step(1, 0.1, t0=0.0 if step == 0 else None, exchange=False)
    for lat in lats:
        lat.stream.synchronize()
    got = torch.cat([lat.local_state() for lat in lats], dim=1).cpu().numpy().reshape(-1)
    assert np.array_equal(bits(got), bits(ref))


@pytest.mark.parametrize("method", ["rk5_i", "rk5_a"])
def test_large_system_adaptive_dopri5(method):
    # compile_sys_openmp's default method is rk5_i (R/odeintr.R:421): dopri5 with the error norm taken over the
    # whole lattice.  Stage algebra runs on the GPU in Boost's order without FMA and the controller's pow() on the
    # host (glibc), so the run follows the oracle step for step; asserted at the stated 10 x rtol tolerance.
    n = 2053
    kw = {} if method == "rk5_i" else {"method": method}
    env = {}
    code = odeintr_b200.compile_sys_openmp("bis", BISTABLE_1D, BISTABLE_PARS, True, sys_dim=n, env=env, **kw)
    assert code.system.gen.stepper.method == method
    ora = bo.build(BISTABLE_1D, BISTABLE_PARS, True, method, sys_dim=n)
    rng = np.random.default_rng(51)
    x0 = rng.uniform(0, 1, n)
    df = env["bis"](x0, 6.0, 0.5)
    ref = ora.integrate_const(x0, 0, 6.0, 0.5)
    assert df.shape == (ref["rows"], n + 1) and np.array_equal(df["Time"].to_numpy(), ref["t"])
    err = np.max(np.abs(cols(df) - ref["rec"]) / (1 + np.abs(ref["rec"])))
    assert err <= 10 * 1e-6, err
    # stronger than required: same image => same glibc pow => the GPU run reproduces the oracle bit for bit
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    st = code.system.stats()
    assert st["accepted"][0] == ref["accepted"] and st["rejected"][0] == ref["rejected"]
    # bistable_at(inic, at) of the reference's example, then the other generated functions
    at = np.array([0.0, 1.0, 10.0 ** 0.5, 10.0])
    df = env["bis_at"](x0, at)
    ref = ora.integrate_times(x0, at, 1.0)
    assert np.max(np.abs(cols(df) - ref["rec"]) / (1 + np.abs(ref["rec"]))) <= 1e-5
    fin = env["bis_no_record"](x0, 4.2, 0.3)
    ref = ora.integrate_adaptive(x0, 0, 4.2, 0.3)
    assert np.max(np.abs(fin - ref["x"])) <= 1e-5
    df = env["bis_adap"](x0, 4.2, 0.3)
    ref = ora.integrate_adaptive(x0, 0, 4.2, 0.3, record=True)
    assert len(df) == ref["rows"] and df["Time"].iloc[-1] == 4.2
    assert np.max(np.abs(df["Time"].to_numpy() - ref["t"])) <= 1e-9
    assert np.max(np.abs(cols(df) - ref["rec"])) <= 1e-5


@pytest.mark.parametrize("case", ["bistable_5003", "chain_two_fields", "fixed_ends", "radius_2", "radius_3",
                                  "bistable_staged"])
@pytest.mark.parametrize("method", ["rk5_i", "rk5_a"])
def test_tiled_dopri5_attempt_bit_exact(case, method, monkeypatch):
    # adaptive dopri5 on 1-D lattices runs one temporally blocked kernel per attempt (odeintr_lattice_dopri5_tiled*):
    # several tiles with a ragged last one, two fields, cells the loop leaves alone, the radius-2 kernel and the
    # general-radius one; "staged" forces the seven-launch path.  All of them must give the oracle's bits: same
    # operation order on the device, the controller's pow() in glibc on both sides.
    pars, glob = BISTABLE_PARS, ""
    if case in ("bistable_5003", "bistable_staged"):
        n, src = 5003, BISTABLE_1D
        x0 = np.random.default_rng(61).uniform(0, 1, n)
        if case == "bistable_staged":
            monkeypatch.setenv("ODEINTR_NO_TILED", "1")
    elif case == "chain_two_fields":
        n, src, pars, glob = 2 * 3001, CHAIN, CHAIN_PARS, CHAIN_GLOBALS
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(n // 2) / (n // 2)), np.zeros(n // 2)])
    elif case == "fixed_ends":
        n = 4100
        src = "for (int i = 1; i < N - 1; ++i) dxdt[i] = D * (x[i - 1] + x[i + 1] - 2.0 * x[i]) + x[i] * (1 - x[i]) * (x[i] - 0.5);"
        x0 = np.random.default_rng(62).uniform(0, 1, n)
    else:
        k = 2 if case == "radius_2" else 3
        n = 4100
        src = ("for (int i = 0; i < N; ++i) dxdt[i] = D * (x[(i + N - %d) %% N] + x[(i + %d) %% N] + x[(i + N - 1) %% N] "
               "+ x[(i + 1) %% N] - 4.0 * x[i]) - 0.2 * x[i];" % (k, k))
        x0 = np.random.default_rng(63).uniform(0, 1, n)
    env = {}
    code = odeintr_b200.compile_sys_openmp("d5", src, pars, True, method=method, sys_dim=n, globals=glob, env=env,
                                           atol=1e-7, rtol=1e-7)
    ora = bo.build(src, pars, True, method, sys_dim=n, globals=glob, atol=1e-7, rtol=1e-7)
    # observations every 0.25: several steps between most of them, more than one inside some steps
    df = env["d5"](x0, 3.0, 0.25)
    ref = ora.integrate_const(x0, 0, 3.0, 0.25)
    assert df.shape == (ref["rows"], n + 1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    st = code.system.stats()
    assert st["accepted"][0] == ref["accepted"] and st["rejected"][0] == ref["rejected"]
    launches = code.system.last_launches()
    tries = int(st["accepted"][0] + st["rejected"][0])
    if case == "bistable_staged":
        assert launches > 6 * tries        # seven stage launches per attempt
    else:
        assert launches < 3 * (tries + len(df)) + 20   # an attempt (and an interpolated observation) is one pass: the warp
                                                       # kernel for the interior + at most two launches for the ends
    at = np.array([0.5, 1.0, 1.0 + 1e-3, 2.75])
    df = env["d5_at"](x0, at, 0.1)
    adv = ora.integrate_const(x0, 0.0, at[0], 0.1, record=False)
    ref = ora.integrate_times(adv["x"], at, 0.1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))


@pytest.mark.parametrize("method", ["rk54", "rk54_a", "rk78", "rk78_a", "rk5"])
def test_large_system_cash_karp_and_fehlberg(method):
    # Cash-Karp 5(4) / Fehlberg 7(8) on a large system: six / thirteen odeintr_lattice_combo launches per step or
    # attempt, tableau-driven on the host (zero entries skipped, Boost's left-to-right sums), controller on the host;
    # rk5 = dopri5 as a plain fixed-step stepper: one temporally blocked attempt per step, k7 handed on as k1
    n = 3001
    env = {}
    code = odeintr_b200.compile_sys_openmp("ck", BISTABLE_1D, BISTABLE_PARS, True, method=method, sys_dim=n, env=env)
    ora = bo.build(BISTABLE_1D, BISTABLE_PARS, True, method, sys_dim=n)
    x0 = np.random.default_rng(52).uniform(0, 1, n)
    df = env["ck"](x0, 3.0, 0.25)
    ref = ora.integrate_const(x0, 0, 3.0, 0.25)
    assert df.shape == (ref["rows"], n + 1) and np.array_equal(df["Time"].to_numpy(), ref["t"])
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    if method.endswith("_a"):
        st = code.system.stats()
        assert st["accepted"][0] == ref["accepted"] and st["rejected"][0] == ref["rejected"]
    at = np.array([0.5, 1.0, 2.2])
    df = env["ck_at"](x0, at, 0.1)
    adv = ora.integrate_const(x0, 0.0, at[0], 0.1, record=False)
    ref = ora.integrate_times(adv["x"], at, 0.1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    fin = env["ck_no_record"](x0, 1.05, 0.1)
    assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0, 1.05, 0.1)["x"]))


def test_2d_lattice_strip_traversal_bit_exact():
    # M a multiple of the block width: the stage kernel walks the lattice in column strips (L1 reuse of the rows
    # above / below); the visiting order must not change a single bit
    m = 512
    env = {}
    odeintr_b200.compile_sys_openmp("b2s", BISTABLE_2D, BISTABLE_PARS, True, method="rk4", sys_dim=m * m, env=env,
                                    halo=0)
    ora = bo.build(BISTABLE_2D, BISTABLE_PARS, True, "rk4", sys_dim=m * m)
    rng = np.random.default_rng(71)
    x0 = rng.integers(0, 2, m * m).astype(np.float64)
    fin = env["b2s_no_record"](x0, 1.0, 0.1)
    ref = ora.integrate_adaptive(x0, 0.0, 1.0, 0.1)
    assert np.array_equal(bits(fin), bits(ref["x"]))


# a hand-written 9-point stencil in the reference's idiom: coordinates from the flat index, neighbours through bound()
NINE_POINT = """
for (int i = 0; i < N; ++i) {
  const int r = i / M, c = i % M;
  dxdt[i] = D * (x[bound(r - 1) * M + c] + x[bound(r + 1) * M + c] + x[r * M + bound(c - 1)] + x[r * M + bound(c + 1)]
                 + 0.5 * (x[bound(r - 1) * M + bound(c - 1)] + x[bound(r + 1) * M + bound(c + 1)]) - 5.0 * x[i]);
}
"""
NINE_POINT_AUTO = NINE_POINT.replace("const int r = i / M, c = i % M;", "const auto r = i / M; const auto c = i % M;")


@pytest.mark.parametrize("case", ["laplace2D_48", "laplace2D_200", "laplace4_doc_96", "nine_point_int", "nine_point_auto",
                                  "inner_cells_only", "reach_3"])
def test_2d_lattice_tiled_rk4_bit_exact(case):
    # M x M lattices step through odeintr_lattice_rk4_tiled2d when the loop body reaches at most 2 rows / columns:
    # ragged tiles (48, 200, 96 are no multiples of the 32 x 64 tile), lattices narrower than a tile, the reference's
    # helpers, hand-written index arithmetic (`int` coordinates take the absolute path, `auto` ones stay symbolic),
    # loops that leave the border alone, and a reach the tile cannot hold (-> stage-per-launch kernels)
    want = 1
    pars = BISTABLE_PARS
    if case == "laplace2D_48":
        m, src = 48, BISTABLE_2D
    elif case == "laplace2D_200":
        m, src = 200, BISTABLE_2D
    elif case == "laplace4_doc_96":
        m, src = 96, BISTABLE_DOC
    elif case == "nine_point_int":
        m, src = 80, NINE_POINT
    elif case == "nine_point_auto":
        m, src = 80, NINE_POINT_AUTO
    elif case == "inner_cells_only":
        m = 72
        src = "for (int i = M + 1; i < N - M - 1; ++i) dxdt[i] = D * (x[i - M] + x[i + M] + x[i - 1] + x[i + 1] - 4.0 * x[i]);"
    else:
        m, want = 64, -1
        src = "for (int i = 0; i < N; ++i) dxdt[i] = D * (x[bound(i / M + 3) * M + i % M] - x[i]);"
    n = m * m
    env = {}
    code = odeintr_b200.compile_sys_openmp("t2", src, pars, True, method="rk4", sys_dim=n, env=env)
    ora = bo.build(src, pars, True, "rk4", sys_dim=n)
    x0 = np.random.default_rng(m).uniform(0, 1, n)
    df = env["t2"](x0, 0.6, 0.1)  # 6 steps, every state observed
    assert code.system.get_halo() == want
    ref = ora.integrate_const(x0, 0, 0.6, 0.1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    fin = env["t2_no_record"](x0, 1.05, 0.1)
    assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0, 1.05, 0.1)["x"]))


def test_documented_bistable_example_with_laplace4():
    # R/odeintr.R:393-413: compile_sys_openmp("bistable", bistable, sys_dim = M * M, pars = c(D = 0.1, a = 1.0, b = 1/2),
    # const = TRUE); bistable_at(inic, 10 ^ (0:3)) -- default method rk5_i, M = 200 in the documentation
    m = 64
    env = {}
    odeintr_b200.compile_sys_openmp("bistable", BISTABLE_DOC, sys_dim=m * m, pars={"D": 0.1, "a": 1.0, "b": 1 / 2},
                                    const=True, env=env)
    ora = bo.build(BISTABLE_DOC, {"D": 0.1, "a": 1.0, "b": 1 / 2}, True, "rk5_i", sys_dim=m * m)
    rng = np.random.default_rng(81)
    inic = rng.integers(0, 2, m * m).astype(np.float64)
    at = 10.0 ** np.arange(0, 3)
    df = env["bistable_at"](inic, at)
    adv = ora.integrate_const(inic, 0.0, at[0], 1.0, record=False)
    ref = ora.integrate_times(adv["x"], at, 1.0)
    assert df.shape == (3, m * m + 1)
    assert np.max(np.abs(cols(df) - ref["rec"])) <= 1e-5


@pytest.mark.parametrize("case", ["bistable", "chain", "wide"])
def test_tiled_rk4_with_declared_halo_bit_exact(case):
    # halo declared -> one temporally blocked kernel per rk4 step (tile + 4*halo cells staged in shared memory);
    # several tiles, a ragged last tile, the periodic seam inside the first / last tile
    if case == "bistable":
        n, src, pars, glob, halo = 5003, BISTABLE_1D, BISTABLE_PARS, "", 1
        x0 = np.random.default_rng(91).integers(0, 2, n).astype(np.float64)
    elif case == "chain":
        n, src, pars, glob, halo = 2 * 3001, CHAIN, CHAIN_PARS, CHAIN_GLOBALS, 1
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(n // 2) / (n // 2)), np.zeros(n // 2)])
    else:  # radius-2 stencil (4th-order Laplacian)
        n, pars, glob, halo = 4100, {"D": 0.05}, "", 2
        src = """
        for (int i = 0; i < N; ++i)
          dxdt[i] = D * (-x[(i + N - 2) % N] + 16.0 * x[(i + N - 1) % N] - 30.0 * x[i] + 16.0 * x[(i + 1) % N] - x[(i + 2) % N]) / 12.0
                    + x[i] * (1 - x[i]) * (x[i] - 0.5);
        """
        x0 = np.random.default_rng(92).uniform(0, 1, n)
    env = {}
    odeintr_b200.compile_sys_openmp("t", src, pars, True, method="rk4", sys_dim=n, globals=glob, env=env, halo=halo)
    ora = bo.build(src, pars, True, "rk4", sys_dim=n, globals=glob)
    df = env["t"](x0, 1.3, 0.1)  # 13 steps, every state observed
    ref = ora.integrate_const(x0, 0, 1.3, 0.1)
    assert df.shape == (ref["rows"], n + 1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    fin = env["t_no_record"](x0, 2.05, 0.1)  # 20 steps + a remainder step
    assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0, 2.05, 0.1)["x"]))


@pytest.mark.parametrize("case", ["fixed_ends", "computed_index", "long_index"])
def test_tiled_rk4_interior_tiles_other_index_forms(case):
    # the interior kernel follows `i + c`, `i - c` and `% N` symbolically; everything else must fall back to the
    # absolute index and still give the staged schedule's bits
    n = 6 * 1024 + 77
    if case == "fixed_ends":  # Dirichlet ends: cells 0 and N-1 are not cells of the loop and keep their value
        src = "for (int i = 1; i < N - 1; ++i) dxdt[i] = D * (x[i - 1] + x[i + 1] - 2.0 * x[i]) + x[i] * (1 - x[i]) * (x[i] - 0.5);"
    elif case == "computed_index":  # i * 1 is an ordinary integer again: generic path of the tile view
        src = "for (int i = 0; i < N; ++i) dxdt[i] = D * (x[(i * 1 + N - 1) % N] + x[(i + 1) % N] - 2.0 * x[i * 1]);"
    else:
        src = "for (size_t i = 0; i < N; ++i) dxdt[i] = D * (x[(i + N - 1) % N] + x[(i + 1) % N] - 2.0 * x[i]) - 0.1 * x[i];"
    pars = {"D": 0.2}
    x0 = np.random.default_rng(93).uniform(0, 1, n)
    env = {}
    odeintr_b200.compile_sys_openmp("t", src, pars, True, method="rk4", sys_dim=n, env=env, halo=1)
    ora = bo.build(src, pars, True, "rk4", sys_dim=n)
    fin = env["t_no_record"](x0, 1.25, 0.1)
    assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0, 1.25, 0.1)["x"]))


def test_tiled_rk4_measured_radius():
    # the probe evaluates the index expressions of every cell: radius 2 for the 4th-order Laplacian, 3 where only a
    # few cells reach that far
    wide = "for (int i = 0; i < N; ++i) dxdt[i] = x[(i + N - 2) % N] + x[(i + 2) % N] - 2.0 * x[i];"
    some = "for (int i = 0; i < N; ++i) dxdt[i] = (i > 2000 && i < 2100 ? x[i + 3] : x[(i + 1) % N]) - x[i];"
    for src, n, want in ((wide, 5000, 2), (some, 5000, 3)):
        env = {}
        code = odeintr_b200.compile_sys_openmp("p", src, BISTABLE_PARS, True, method="rk4", sys_dim=n, env=env)
        x0 = np.random.default_rng(94).uniform(0, 1, n)
        fin = env["p_no_record"](x0, 0.5, 0.1)
        assert code.system.get_halo() == want
        ora = bo.build(src, BISTABLE_PARS, True, "rk4", sys_dim=n)
        assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0, 0.5, 0.1)["x"]))


def test_tiled_rk4_state_dependent_reach_falls_back_to_the_staged_kernels():
    # the probe (which reads 1.0 everywhere) sees radius 1; the run reads 5 cells away wherever x > 1.5.  The
    # per-access check catches it and the call is repeated with the stage-per-launch kernels: same bits as the oracle
    src = "for (int i = 0; i < N; ++i) dxdt[i] = x[(i + (x[i] > 1.5 ? 5 : 1)) % N] - x[i];"
    n = 5000
    env = {}
    code = odeintr_b200.compile_sys_openmp("f", src, method="rk4", sys_dim=n, env=env)
    x0 = np.random.default_rng(95).uniform(0, 2, n)
    df = env["f"](x0, 1.0, 0.1)
    assert code.system.get_halo() == -1  # switched off for this system from now on
    ora = bo.build(src, None, False, "rk4", sys_dim=n)
    ref = ora.integrate_const(x0, 0, 1.0, 0.1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))


def test_large_system_rejects_strided_targets():
    with pytest.raises(odeintr_b200.OdeintrError, match="loop variable \\+ constant"):
        odeintr_b200.compile_sys_openmp("s", "for (int i = 0; i < N / 2; ++i) { dxdt[2 * i] = x[2 * i + 1]; dxdt[2 * i + 1] = -x[2 * i]; }",
                                        method="rk4", sys_dim=64)


def test_tiled_rk4_wide_radius_and_declared_radius_variants():
    # radius 8 (margins of 32 cells, the TMA windows shift by a multiple of 16 bytes), a declared radius larger than
    # the stencil needs, and a 2-D stencil declared with radius 2 (the _h2 kernel on a radius-1 Laplacian)
    src8 = "for (int i = 0; i < N; ++i) dxdt[i] = D * (x[(i + N - 8) % N] + x[(i + 8) % N] + x[(i + 1) % N] - 3.0 * x[i]);"
    for src, n, halo, want in ((src8, 9001, None, 8), (BISTABLE_1D, 7000, 3, 3), (BISTABLE_2D, 96 * 96, 2, 2),
                               (BISTABLE_2D, 96 * 96, 1, 1)):
        env = {}
        code = odeintr_b200.compile_sys_openmp("v", src, BISTABLE_PARS, True, method="rk4", sys_dim=n, env=env, halo=halo)
        x0 = np.random.default_rng(n).uniform(0, 1, n)
        fin = env["v_no_record"](x0, 0.7, 0.1)
        assert code.system.get_halo() == want
        ora = bo.build(src, BISTABLE_PARS, True, "rk4", sys_dim=n)
        assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0, 0.7, 0.1)["x"]))


def test_sharded_tiled_slabs_with_fixed_ends_emulated_ranks():
    # non-periodic loop (cells 0 and N-1 keep their value) split into three slabs: the slabs at the lattice ends hold
    # cells the loop does not write, so their first / last tiles take the general tiled kernel
    import torch

    from odeintr_b200.distributed import ShardedLattice

    n_cells, world = 7001, 3
    src = "for (int i = 1; i < N - 1; ++i) dxdt[i] = D * (x[i - 1] + x[i + 1] - 2.0 * x[i]) + x[i] * (1 - x[i]) * (x[i] - 0.5);"
    x0 = np.random.default_rng(97).uniform(0, 1, n_cells)
    build = lambda name: odeintr_b200.compile_sys_openmp(name, src, {"D": 0.2}, True, method="rk4", sys_dim=n_cells).system
    ora = bo.build(src, {"D": 0.2}, True, "rk4", sys_dim=n_cells)
    rec = ora.integrate_const(x0, 0, 1.25, 0.1)  # 12 whole steps of 0.1 (no remainder step)
    assert rec["rows"] == 13
    ref = rec["rec"][-1]
    lats = []
    for r in range(world):
        lat = ShardedLattice(build("r%d" % r), n_cells, fields=1, rank=r, world=world, native=False)
        lat.set_state_from(lambda f, cells: x0[cells])
        lats.append(lat)
    g, pad = lats[0].ghost, lats[0].pad
    for step in range(12):
        for lat in lats:
            lat.stream.synchronize()
        for r, lat in enumerate(lats):
            left, right = lats[(r - 1) % world], lats[(r + 1) % world]
            lat.x[:, pad - g:pad] = left.x[:, pad + left.n_local - g:pad + left.n_local]
            lat.x[:, pad + lat.n_local:pad + lat.n_local + g] = right.x[:, pad:pad + g]
        torch.cuda.synchronize()
        for lat in lats:
            lat.step(1, 0.1, t0=0.0 if step == 0 else None, exchange=False)
    for lat in lats:
        lat.stream.synchronize()
    got = torch.cat([lat.local_state() for lat in lats], dim=1).cpu().numpy().reshape(-1)
    bad = np.nonzero(bits(got) != bits(ref))[0]
    assert bad.size == 0, (bad.size, bad[:12], bad[-12:])


@pytest.mark.parametrize("method", ["rk4", "rk5_i"])
def test_tiny_lattices(method):
    # lattices shorter than a tile's margins: 1-D with 5 and 9 cells, 2-D with M = 3 and M = 8
    for src, n in ((BISTABLE_1D, 5), (BISTABLE_1D, 9), (BISTABLE_1D, 40), (BISTABLE_2D, 9), (BISTABLE_2D, 64)):
        env = {}
        odeintr_b200.compile_sys_openmp("tiny", src, BISTABLE_PARS, True, method=method, sys_dim=n, env=env)
        ora = bo.build(src, BISTABLE_PARS, True, method, sys_dim=n)
        x0 = np.random.default_rng(n).uniform(0, 1, n)
        df = env["tiny"](x0, 2.0, 0.25)
        ref = ora.integrate_const(x0, 0, 2.0, 0.25)
        assert df.shape == (ref["rows"], n + 1)
        assert np.array_equal(bits(cols(df)), bits(ref["rec"])), (method, n)


@pytest.mark.parametrize("method", ["rk4", "rk5_i"])
@pytest.mark.parametrize("case", ["clamped_ends", "user_template", "index_as_number"])
def test_snippets_that_use_the_loop_variable_as_a_plain_int(case, method):
    # clamped (zero-flux) ends through std::min / std::max keep the tiled kernels (mixed-type overloads, absolute-index
    # path of the views); a snippet whose helper template cannot take the symbolic index is built with the
    # stage-per-launch kernels only.  Same bits as the oracle either way.
    n = 4500
    if case == "clamped_ends":
        src, glob = "for (int i = 0; i < N; ++i) dxdt[i] = D * (x[std::min(i + 1, (int)N - 1)] + x[std::max(i - 1, 0)] - 2.0 * x[i]) - 0.1 * x[i];", ""
    elif case == "index_as_number":
        # the wrapped neighbour index also enters the arithmetic as a number, and is shifted after the wrap
        src, glob = ("for (int i = 0; i < N; ++i) dxdt[i] = D * (x[(i + 1) % N] - x[i]) + 1e-4 * ((i + 3) % N) "
                     "- 0.1 * x[((i + N - 1) % N + 1) % N] + (i > 100 ? 1e-3 : 0.0);"), ""
    else:
        glob = "template <class T> T pick(T a, T b) { return a < b ? a : b; }"
        src = "for (int i = 0; i < N; ++i) dxdt[i] = D * (x[pick(i + 1, (int)N - 1)] - x[i]) - 0.1 * x[i];"
    env = {}
    code = odeintr_b200.compile_sys_openmp("s", src, {"D": 0.3}, True, method=method, sys_dim=n, globals=glob, env=env)
    ora = bo.build(src, {"D": 0.3}, True, method, sys_dim=n, globals=glob)
    x0 = np.random.default_rng(98).uniform(0, 1, n)
    df = env["s"](x0, 1.5, 0.25)
    ref = ora.integrate_const(x0, 0, 1.5, 0.25)
    assert df.shape == (ref["rows"], n + 1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    if method == "rk4":
        assert code.system.get_halo() == (-1 if case == "user_template" else 1)


def test_tiled_rk4_interior_kernel_flags_a_wide_stencil_too():
    # only interior tiles see the far read: cells next to the ends stay within reach
    src = "for (int i = 0; i < N; ++i) dxdt[i] = (i > 2000 && i < 2100 ? x[i + 3] : x[(i + 1) % N]) - x[i];"
    env = {}
    odeintr_b200.compile_sys_openmp("w", src, method="rk4", sys_dim=5000, env=env, halo=1)
    with pytest.raises(odeintr_b200.OdeintrError, match="further than the declared halo"):
        env["w"](np.ones(5000), 1.0, 0.1)


def test_tiled_rk4_detects_a_stencil_wider_than_declared():
    src = "for (int i = 0; i < N; ++i) dxdt[i] = x[(i + 2) % N] - x[i];"
    env = {}
    odeintr_b200.compile_sys_openmp("w", src, method="rk4", sys_dim=3000, env=env, halo=1)
    with pytest.raises(odeintr_b200.OdeintrError, match="further than the declared halo"):
        env["w"](np.ones(3000), 1.0, 0.1)


@pytest.mark.parametrize("case,spe,steps", [("bistable", 8, (19, 6)), ("bistable", 1, (5, 3)), ("bistable", 3, (7, 3)),
                                            ("chain", 8, (16, 9)), ("chain", 2, (5, 4)), ("fixed_ends", 4, (9, 4))])
def test_native_shard_run_overlapped_exchange_matches_unsharded(case, spe, steps, monkeypatch):
    # odeintr_lattice_shard_run sends the halo exchange down a second stream: the last step of a round computes its
    # edge cells first, the exchange takes them while the middle of the slab is computed, the first step of the next
    # round waits for the ghost cells only before its own edges.  Whatever the round length, the number of steps per
    # call and where a call ends within a round, the slab must carry the bits of the unsharded run; the serial
    # schedule (ODEINTR_NO_OVERLAP) is checked beside it.
    import torch

    from odeintr_b200.distributed import ShardedLattice

    if case == "chain":
        n_cells, fields, src, pars, glob = 2 * 5003, 2, CHAIN, CHAIN_PARS, CHAIN_GLOBALS
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(n_cells) / n_cells), 0.2 * np.cos(2 * np.pi * 5 * np.arange(n_cells) / n_cells)])
    else:
        n_cells, fields, pars, glob = 20011, 1, BISTABLE_PARS, ""
        src = BISTABLE_1D if case == "bistable" else \
            "for (int i = 1; i < N - 1; ++i) dxdt[i] = D * (x[i - 1] + x[i + 1] - 2.0 * x[i]) + a * x[i] * (1 - x[i]) * (x[i] - b);"
        x0 = np.random.default_rng(43).uniform(0, 1, n_cells)
    build = lambda name: odeintr_b200.compile_sys_openmp(name, src, pars, True, method="rk4", sys_dim=fields * n_cells,
                                                         globals=glob).system
    total = sum(steps)
    ref = build("ref").no_record(x0, total * 0.1, 0.1)
    for overlap in (True, False):
        if not overlap:
            monkeypatch.setenv("ODEINTR_NO_OVERLAP", "1")
        lat = ShardedLattice(build("sh"), n_cells, fields=fields, halo=1, rank=0, world=1, native=True, steps_per_exchange=spe)
        lat.set_state_from(lambda f, cells: x0[f * n_cells + cells])
        lat.step(steps[0], 0.1, t0=0.0)
        lat.step(steps[1], 0.1)
        torch.cuda.synchronize()
        got = lat.gather_state().cpu().numpy().reshape(-1)
        assert np.array_equal(bits(got), bits(ref)), (case, spe, overlap)


@pytest.mark.parametrize("method", ["rk5_i", "rk5_a"])
@pytest.mark.parametrize("case", ["bistable", "chain", "fixed_ends"])
def test_sharded_adaptive_dopri5_single_rank_matches_unsharded(case, method):
    # compile_sys_openmp's default method on a slab (odeintr_lattice_shard + the ordinary integrate calls): the attempt
    # kernel addresses the slab with its ghost cells, the derivative alone comes from the sharded stage kernel, ghost
    # cells are refreshed after every accepted step.  With one rank the ring closes on the slab itself; the result must
    # be the unsharded run's (hence the oracle's) bit for bit, step counts included.
    import torch

    from odeintr_b200.distributed import ShardedLattice

    if case == "chain":
        n_cells, fields, src, pars, glob = 5003, 2, CHAIN, CHAIN_PARS, CHAIN_GLOBALS
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(n_cells) / n_cells), 0.2 * np.cos(2 * np.pi * 5 * np.arange(n_cells) / n_cells)])
    else:
        n_cells, fields, pars, glob = 9001, 1, BISTABLE_PARS, ""
        src = BISTABLE_1D if case == "bistable" else \
            "for (int i = 1; i < N - 1; ++i) dxdt[i] = D * (x[i - 1] + x[i + 1] - 2.0 * x[i]) + a * x[i] * (1 - x[i]) * (x[i] - b);"
        x0 = np.random.default_rng(45).uniform(0, 1, n_cells)
    kw = dict(method=method, sys_dim=fields * n_cells, globals=glob, atol=1e-7, rtol=1e-7)
    env = {}
    ref_sys = odeintr_b200.compile_sys_openmp("ref", src, pars, True, env=env, **kw).system
    at = np.array([0.0, 0.5, 1.0 + 1e-3, 2.75])
    ref = env["ref_at"](x0, at, 0.1)
    ref_stats = ref_sys.stats()
    lat = ShardedLattice(odeintr_b200.compile_sys_openmp("sh", src, pars, True, **kw).system, n_cells, fields=fields,
                         halo=1, rank=0, world=1, native=True)
    lat.set_state_from(lambda f, cells: x0[f * n_cells + cells])
    t, x = lat.integrate_times(at, 0.1)
    assert np.array_equal(t, at) and x.shape == (4, fields, n_cells)
    assert np.array_equal(bits(x.reshape(4, -1)), bits(cols(ref)))
    acc, rej = lat.stats()
    assert acc == ref_stats["accepted"][0] and rej == ref_stats["rejected"][0]
    # the slab keeps the final state: continue with NAME_no_record / NAME on it
    fin_ref = env["ref_no_record"](cols(ref)[-1], 0.7, 0.05, 2.75)
    lat.no_record(0.7, 0.05, start=2.75)
    torch.cuda.synchronize()
    assert np.array_equal(bits(lat.gather_state().cpu().numpy().reshape(-1)), bits(fin_ref))
    rec_ref = env["ref"](fin_ref, 1.0, 0.25, 3.45)
    t, x = lat.integrate_const(1.0, 0.25, start=3.45)
    assert np.array_equal(t, rec_ref["Time"].to_numpy()) and np.array_equal(bits(x.reshape(len(t), -1)), bits(cols(rec_ref)))


MEAN_FIELD = """
double m = 0.0;
for (int i = 0; i < N; ++i) m += x[i];
m /= N;
#pragma omp parallel for
for (int i = 0; i < N; ++i) dxdt[i] = -x[i] + K * (m - x[i]) * x[i];
dxdt[0] += 0.1 * std::sin(t);
"""
NESTED_2D = """
for (int i = 0; i < M; ++i)
  for (int j = 0; j < M; ++j)
    dxdt[i * M + j] = D * laplace2D(x, i, j) + x[i * M + j] * (1 - x[i * M + j]) * (x[i * M + j] - 0.5);
"""


@pytest.mark.parametrize("method", ["rk4", "euler", "rk5_i", "rk5_a", "rk54_a", "rk78", "rk5", "bs", "bsd"])
@pytest.mark.parametrize("case", ["mean_field", "nested_2d", "straight_line"])
def test_general_snippets_take_the_serial_rhs_path(case, method):
    # compile_sys_openmp pastes ANY C++ into sys() (inst/templates/compile_sys_openmp_template.cpp:39-43).  Snippets
    # that are not loops over the cells -- a reduction ahead of the loop (mean-field coupling), nested i / j loops,
    # statements outside any loop, no loop at all -- cannot be dealt out one cell per thread; they run on the general
    # path: the snippet verbatim in one thread (the reference's own evaluation order), the steppers' vector algebra in
    # parallel on stored stage derivatives.  Same bits as the oracle, for every method of the large-system table.
    if case == "mean_field":
        n, src, pars = 777, MEAN_FIELD, {"K": 0.5}
    elif case == "nested_2d":
        n, src, pars = 24 * 24, NESTED_2D, {"D": 0.1}
    else:
        n, src, pars = 2, "dxdt[0] = x[1]; dxdt[1] = -x[0] - 0.1 * x[1];", None
    env = {}
    code = odeintr_b200.compile_sys_openmp("g", src, pars, True, method=method, sys_dim=n, env=env)
    assert "odeintr_lattice_serial_rhs" in code
    ora = bo.build(src, pars, True, method, sys_dim=n)
    x0 = np.random.default_rng(n).uniform(0, 1, n)
    df = env["g"](x0, 2.0, 0.25)
    ref = ora.integrate_const(x0, 0, 2.0, 0.25)
    assert df.shape == (ref["rows"], n + 1) and np.array_equal(df["Time"].to_numpy(), ref["t"])
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    at = np.array([0.5, 1.0, 2.2])
    df = env["g_at"](x0, at, 0.1)
    adv = ora.integrate_const(x0, 0.0, at[0], 0.1, record=False)
    ref = ora.integrate_times(adv["x"], at, 0.1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    fin = env["g_no_record"](x0, 1.05, 0.1)
    assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0, 1.05, 0.1)["x"]))


@pytest.mark.parametrize("case", ["bistable", "chain", "fixed_ends", "radius_2"])
def test_two_rk4_steps_per_pass_bit_exact(case, monkeypatch):
    # lattice_step_pair: the warp kernel takes two rk4 steps per pass (the state between them never reaches HBM), the
    # cells at the ends of the loop range / the periodic seam go through the general kernel twice via an intermediate
    # array.  Odd step counts leave a single step at the end; observers between steps break the pairs up; loops that
    # leave cells alone need those cells in the intermediate array too.  Forced on here (by default it is used where a
    # single step is HBM-bound: several fields); every case must give the oracle's bits.
    monkeypatch.setenv("ODEINTR_FUSE", "1")
    pars, glob = BISTABLE_PARS, ""
    if case == "chain":
        L = 20010
        n, src, pars, glob = 2 * L, CHAIN, CHAIN_PARS, CHAIN_GLOBALS
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(L) / L), 0.1 * np.cos(2 * np.pi * 3 * np.arange(L) / L)])
    elif case == "bistable":
        n, src = 30011, BISTABLE_1D
        x0 = np.random.default_rng(101).uniform(0, 1, n)
    elif case == "fixed_ends":
        n = 30011
        src = "for (int i = 1; i < N - 1; ++i) dxdt[i] = D * (x[i - 1] + x[i + 1] - 2.0 * x[i]) + a * x[i] * (1 - x[i]) * (x[i] - b);"
        x0 = np.random.default_rng(102).uniform(0, 1, n)
    else:
        n = 30011
        src = ("for (int i = 0; i < N; ++i) dxdt[i] = D * (x[(i + N - 2) % N] + x[(i + 2) % N] + x[(i + N - 1) % N] "
               "+ x[(i + 1) % N] - 4.0 * x[i]) - 0.2 * x[i];")
        x0 = np.random.default_rng(103).uniform(0, 1, n)
    env = {}
    code = odeintr_b200.compile_sys_openmp("f2", src, pars, True, method="rk4", sys_dim=n, globals=glob, env=env)
    ora = bo.build(src, pars, True, "rk4", sys_dim=n, globals=glob)
    fin = env["f2_no_record"](x0, 0.7, 0.1)  # 7 steps: three pairs and a single step
    launches = code.system.last_launches()
    assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0.0, 0.7, 0.1)["x"]))
    assert launches <= 3 * 3 + 3, launches  # a pair costs at most three launches (two for the ends, one for the interior)
    # observer every third step (obs_stride through the C-ABI): pairs inside a segment, a single step left over
    s = code.system
    s.set_state(x0)
    s.reset_observer()
    from odeintr_b200 import _abi
    assert s._lib.odeintr_integrate_const(s._h, 0.0, 0.9, 0.1, 3, 1) == 0, _abi.last_error()
    rec = s.get_output()
    ref = ora.integrate_const(x0, 0.0, 0.9, 0.1)
    assert np.array_equal(bits(cols(rec)), bits(ref["rec"][::3]))


@pytest.mark.parametrize("case", ["laplace2D_330", "nine_point_128_short_segments", "laplace4_doc_200"])
def test_2d_lattice_streaming_rk4_bit_exact(case, monkeypatch):
    # M x M lattices with radius 1, even M >= 64 and a loop over all cells step through the streaming kernel
    # (lattice_stream2d.cuh): strips of 56 columns (ragged last strip, columns wrapping across the seam), row segments
    # (ragged last one, 8 rows of pipeline fill wrapping across the seam), the 9-point stencil's diagonal neighbours.
    # Must give the oracle's bits, and the block-tiled kernel (ODEINTR_NO_STREAM2D) must too.
    pars = BISTABLE_PARS
    if case == "laplace2D_330":
        m, src = 330, BISTABLE_2D          # 6 strips (5.9), 2 row segments (256 + 74)
    elif case == "nine_point_128_short_segments":
        m, src = 128, NINE_POINT_AUTO
        monkeypatch.setenv("ODEINTR_STREAM2_ROWS", "24")   # 6 segments, the last one ragged
    else:
        m, src = 200, BISTABLE_DOC
    ora = bo.build(src, pars, True, "rk4", sys_dim=m * m)
    x0 = np.random.default_rng(m).uniform(0, 1, m * m)
    ref = ora.integrate_adaptive(x0, 0.0, 0.7, 0.1)["x"]
    for stream in (True, False):
        if not stream:
            monkeypatch.setenv("ODEINTR_NO_STREAM2D", "1")
        env = {}
        code = odeintr_b200.compile_sys_openmp("s2", src, pars, True, method="rk4", sys_dim=m * m, env=env)
        fin = env["s2_no_record"](x0, 0.7, 0.1)
        assert code.system.get_halo() == 1
        assert np.array_equal(bits(fin), bits(ref)), (case, stream)


FIXED_ENDS = """
for (int i = 1; i < N - 1; ++i)
  dxdt[i] = D * (x[i - 1] + x[i + 1] - 2.0 * x[i]) + a * x[i] * (1 - x[i]) * (x[i] - b);
"""


@pytest.mark.parametrize("method", ["bs", "bsd"])
@pytest.mark.parametrize("case", ["periodic", "fixed_ends", "chain"])
def test_large_system_bulirsch_stoer(case, method):
    # bs / bsd on large systems (R/utils.R:57-58 with openmp_range_algebra): modified-midpoint runs, extrapolation
    # tables, finite differences and the dense-output polynomial as odeintr_lattice_combo launches in Boost's order, the
    # order / step-size controller on the host (glibc pow, like the oracle): bit-exact, step counts included.  The
    # fixed-ends case has entries no loop writes: the extrapolation (1 + c) x - c x does not reproduce them exactly, so
    # the combos cover every entry.
    if case == "chain":
        n, src, pars, glob = 2 * 700, CHAIN, CHAIN_PARS, CHAIN_GLOBALS
    else:
        n, src, pars, glob = 2003, (BISTABLE_1D if case == "periodic" else FIXED_ENDS), BISTABLE_PARS, ""
    env = {}
    code = odeintr_b200.compile_sys_openmp("bsl", src, pars, True, method=method, sys_dim=n, env=env, globals=glob)
    ora = bo.build(src, pars, True, method, sys_dim=n, globals=glob)
    x0 = np.random.default_rng(77).uniform(0.1, 1, n)
    df = env["bsl"](x0, 3.0, 0.25)
    ref = ora.integrate_const(x0, 0, 3.0, 0.25)
    assert df.shape == (ref["rows"], n + 1) and np.array_equal(df["Time"].to_numpy(), ref["t"])
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    st = code.system.stats()
    assert st["accepted"][0] == ref["accepted"] and st["rejected"][0] == ref["rejected"]
    at = np.array([0.5, 1.0, 2.2, 7.0])
    df = env["bsl_at"](x0, at, 0.1)
    adv = ora.integrate_const(x0, 0.0, at[0], 0.1, record=False)
    ref = ora.integrate_times(adv["x"], at, 0.1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    fin = env["bsl_no_record"](x0, 4.05, 0.1)
    assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0, 4.05, 0.1)["x"]))


@pytest.mark.parametrize("method", ["rk4", "rk5_i"])
def test_2d_lattice_streaming_kernels_several_tasks_per_warp(method, monkeypatch):
    # On the benchmark lattices (8192 x 8192: ~3500 tasks over 1184 / 2368 resident warps) every warp of the persistent
    # streaming kernels walks through several (strip, row segment) tasks, re-priming its register pipeline for each.  At
    # test sizes there are fewer tasks than warps, so the grid is capped at two blocks here: 6 strips x 14 segments = 84
    # tasks over 16 warps.  Oracle's bits, for K4s2 and K4ds2.
    m = 330
    monkeypatch.setenv("ODEINTR_STREAM_MAX_BLOCKS", "2")
    monkeypatch.setenv("ODEINTR_STREAM2_ROWS", "24")
    monkeypatch.setenv("ODEINTR_STREAM5_ROWS", "24")
    env = {}
    code = odeintr_b200.compile_sys_openmp("mt", BISTABLE_2D, BISTABLE_PARS, True, method=method, sys_dim=m * m, env=env)
    ora = bo.build(BISTABLE_2D, BISTABLE_PARS, True, method, sys_dim=m * m)
    x0 = np.random.default_rng(m + 1).uniform(0, 1, m * m)
    fin = env["mt_no_record"](x0, 0.7, 0.1)
    assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0.0, 0.7, 0.1)["x"]))
    if method == "rk5_i":
        st = code.system.stats()
        assert code.system.last_launches() <= st["accepted"][0] + st["rejected"][0] + 3  # the streaming kernel took the attempts


@pytest.mark.parametrize("method", ["rk5_i", "rk5_a", "rk5"])
@pytest.mark.parametrize("case", ["laplace2D_128", "documented_192", "laplace2D_66", "staged_128"])
def test_2d_lattice_streaming_dopri5_bit_exact(case, method, monkeypatch):
    # The reference's documented large-system example (R/odeintr.R:393-421: an M x M bistable lattice, default method
    # rk5_i) through K4ds2: one attempted dopri5 step per launch, warps streaming down strips of 64 columns with the
    # six neighbour-dependent stages as a register pipeline and partial sums in place of stored stage derivatives.
    # Same bits as the oracle (controller on the host: glibc pow), same accepted / rejected counts; M = 192 and 66 are
    # not multiples of the 52 columns a strip yields (ragged last strip, wrap across the seam); the staged case pins
    # the seven-launch path beside it.
    m = {"laplace2D_128": 128, "documented_192": 192, "laplace2D_66": 66, "staged_128": 128}[case]
    src = BISTABLE_DOC if case == "documented_192" else BISTABLE_2D
    if case == "staged_128":
        monkeypatch.setenv("ODEINTR_NO_STREAM2D", "1")
    n = m * m
    env = {}
    code = odeintr_b200.compile_sys_openmp("d2", src, BISTABLE_PARS, True, method=method, sys_dim=n, env=env)
    ora = bo.build(src, BISTABLE_PARS, True, method, sys_dim=n)
    x0 = np.random.default_rng(m).uniform(0, 1, n)
    df = env["d2"](x0, 3.0, 0.25)
    ref = ora.integrate_const(x0, 0, 3.0, 0.25)
    assert df.shape == (ref["rows"], n + 1) and np.array_equal(df["Time"].to_numpy(), ref["t"])
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    if method != "rk5":
        st = code.system.stats()
        assert st["accepted"][0] == ref["accepted"] and st["rejected"][0] == ref["rejected"]
        tries = ref["accepted"] + ref["rejected"]
        launches = code.system.last_launches()
        if case == "staged_128":
            assert launches >= 7 * tries
        else:  # one launch per attempt (+ the dense-output evaluations and the first derivative)
            assert launches <= tries + ref["rows"] + 2, (launches, tries)
    at = np.array([0.0, 1.0, 10.0 ** 0.5, 10.0])  # bistable_at(inic, 10^(0:3))-style call
    df = env["d2_at"](x0, at, 0.1)
    ref = ora.integrate_times(x0, at, 0.1)
    assert np.array_equal(bits(cols(df)), bits(ref["rec"]))
    fin = env["d2_no_record"](x0, 2.05, 0.1)
    assert np.array_equal(bits(fin), bits(ora.integrate_adaptive(x0, 0, 2.05, 0.1)["x"]))
