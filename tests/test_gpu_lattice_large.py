"""GPU parity of the large-system kernels at sizes where every block does more than one tile and the grid more than
one wave: the persistent loops of the temporally blocked rk4 kernels (prefetch into the other
buffer, mbarrier parity flip, output staging), the tiled dopri5 attempt and the 2-D tiled step, against the oracle
running the same loops on the CPU -- and BASELINE.json configs[4] at its full size (2^28 cells).  Fixed-step
arithmetic in Boost's order without FMA -> bit-exact."""
import numpy as np
import pytest

import odeintr_b200
from oracle import build_oracle as bo
from systems import BISTABLE_1D, BISTABLE_2D, BISTABLE_PARS, CHAIN, CHAIN_GLOBALS, CHAIN_PARS

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def sm_count():
    import torch

    return torch.cuda.get_device_properties(0).multi_processor_count


@pytest.mark.parametrize("case", ["one_field", "two_fields", "one_field_odd_start", "one_field_two_steps_per_pass"])
def test_persistent_tiled_rk4_many_tiles_per_block_bit_exact(case, monkeypatch):
    # 2^21 + 6 cells: > 2000 tiles of 1024 cells (> 8000 warp tiles) on at most 148 x 4 resident blocks, so every
    # persistent block walks through several tiles; the ragged end and the periodic seam go to the general kernel
    if case == "two_fields":
        L = (1 << 20) + 2  # even distance between the fields: TMA-aligned windows
        n, src, pars, glob = 2 * L, CHAIN, CHAIN_PARS, CHAIN_GLOBALS
        x0 = np.concatenate([np.sin(2 * np.pi * 8 * np.arange(L) / L), 0.1 * np.cos(2 * np.pi * 3 * np.arange(L) / L)])
    elif case in ("one_field", "one_field_two_steps_per_pass"):
        if case != "one_field":
            monkeypatch.setenv("ODEINTR_FUSE", "1")  # (the default for several fields: "two_fields" above)
        n, src, pars, glob = (1 << 21) + 6, BISTABLE_1D, BISTABLE_PARS, ""
        x0 = np.random.default_rng(81).uniform(0, 1, n)
    else:
        # a loop that starts at an odd cell: windows that are not 16-byte aligned take the non-TMA interior kernel
        n, pars, glob = (1 << 21) + 5, BISTABLE_PARS, ""
        src = "for (int i = 1; i < N - 1; ++i) dxdt[i] = D * (x[i - 1] + x[i + 1] - 2.0 * x[i]) + a * x[i] * (1 - x[i]) * (x[i] - b);"
        x0 = np.random.default_rng(82).uniform(0, 1, n)
    env = {}
    code = odeintr_b200.compile_sys_openmp("big", src, pars, True, method="rk4", sys_dim=n, globals=glob, env=env)
    assert code.system.get_halo() == 1
    ora = bo.build(src, pars, True, "rk4", sys_dim=n, globals=glob)
    fin = env["big_no_record"](x0, 0.7, 0.1)  # 7 steps: the buffers trade places an odd number of times
    ref = ora.integrate_adaptive(x0, 0.0, 0.7, 0.1)
    assert np.array_equal(bits(fin), bits(ref["x"]))
    # more tiles than resident blocks, whatever kernel took the interior
    assert n // 1024 > 3 * 4 * sm_count()


def test_config4_full_size_two_to_the_28_cells_bit_exact():
    # BASELINE.json configs[4] at its full size: 2^28 cells, 5 rk4 steps, all 2 GiB of the final state compared with
    # the oracle bit for bit (the oracle runs the user's `#pragma omp parallel for` loop and the stepper's algebra on
    # all host cores, ~1 minute)
    n = 1 << 28
    env = {}
    odeintr_b200.compile_sys_openmp("c4", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n, env=env)
    ora = bo.build(BISTABLE_1D, BISTABLE_PARS, True, "rk4", sys_dim=n)
    # deterministic, aperiodic initial pattern without a 2 GiB random draw: a few incommensurate waves
    i = np.arange(n, dtype=np.float64)
    x0 = 0.5 + 0.25 * np.sin(i * 0.001) + 0.2 * np.sin(i * 0.0371 + 1.0) + 0.05 * np.sin(i * 1.7)
    del i
    fin = env["c4_no_record"](x0, 0.05, 0.01)
    ref = ora.integrate_adaptive(x0, 0.0, 0.05, 0.01)
    assert np.array_equal(bits(fin), bits(ref["x"]))
    assert not np.array_equal(bits(fin), bits(x0))


@pytest.mark.parametrize("method", ["rk5_i", "rk5_a"])
def test_tiled_dopri5_more_tiles_than_resident_blocks_bit_exact(method):
    # 2^20 + 3 cells = 1025 tiles of the dopri5 attempt kernel against at most 148 x 3 resident blocks: several waves,
    # the max-norm reduction across all of them, the fused dense-output pass
    n = (1 << 20) + 3
    env = {}
    code = odeintr_b200.compile_sys_openmp("d5big", BISTABLE_1D, BISTABLE_PARS, True, method=method, sys_dim=n, env=env,
                                           atol=1e-7, rtol=1e-7)
    ora = bo.build(BISTABLE_1D, BISTABLE_PARS, True, method, sys_dim=n, atol=1e-7, rtol=1e-7)
    x0 = np.random.default_rng(83).uniform(0, 1, n)
    at = np.array([0.0, 0.5, 1.25])
    df = env["d5big_at"](x0, at, 0.1)
    ref = ora.integrate_times(x0, at, 0.1)
    assert np.array_equal(bits(df.iloc[:, 1:].to_numpy()), bits(ref["rec"]))
    st = code.system.stats()
    assert st["accepted"][0] == ref["accepted"] and st["rejected"][0] == ref["rejected"]
    assert n // 1024 > 2 * 3 * sm_count() or sm_count() > 170


def test_tiled2d_more_tiles_than_resident_blocks_bit_exact():
    # 1100 x 1100 lattice through laplace2D: 35 x 18 = 630 tiles of 32 x 64 cells (> 148 x 3 resident blocks), ragged
    # in both directions
    m = 1100
    env = {}
    code = odeintr_b200.compile_sys_openmp("b2big", BISTABLE_2D, BISTABLE_PARS, True, method="rk4", sys_dim=m * m, env=env)
    assert code.system.get_halo() == 1
    ora = bo.build(BISTABLE_2D, BISTABLE_PARS, True, "rk4", sys_dim=m * m)
    x0 = np.random.default_rng(84).integers(0, 2, m * m).astype(np.float64)
    fin = env["b2big_no_record"](x0, 0.5, 0.1)
    ref = ora.integrate_adaptive(x0, 0.0, 0.5, 0.1)
    assert np.array_equal(bits(fin), bits(ref["x"]))
