"""Host-side multi-rank logic on CPU (gloo, world_size 2): instance sharding, the periodic halo exchange, and
the one-exchange-per-step wide-halo rk4 scheme that odeintr_lattice_shard_step implements on the GPU — here
emulated in numpy on each rank's slab and compared bit for bit with the oracle on the whole lattice."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from odeintr_b200.distributed import exchange_halos, shard_range


def test_shard_ranges_partition_everything():
    for total in (1, 7, 100, 100_000_000, 2 ** 28 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rhs(x_window, D=0.1, a=1.0, b=0.5):
    """BISTABLE_1D on the interior of a window (needs one valid cell on each side); same operation order."""
    c, l, r = x_window[1:-1], x_window[:-2], x_window[2:]
    return D * (l + r - 2.0 * c) + a * c * (1 - c) * (c - b)


def _sharded_rk4_step(x, ghost, dt):
    """One rk4 step on a slab with `ghost` = 4 valid ghost cells per side, shrinking window per stage,
    running-sum accumulation in Boost's order (what the fused stage kernels do)."""
    half, b1, b2 = 1.0 / 2.0, 1.0 / 6.0, 1.0 / 3.0
    n = x.shape[0]
    A, B, acc = x.copy(), x.copy(), x.copy()
    k = _rhs(x[0:n]);           w = slice(1, n - 1)
    A[w] = x[w] + (half * dt) * k; acc[w] = x[w] + (b1 * dt) * k
    k = _rhs(A[1:n - 1]);       w = slice(2, n - 2)
    B[w] = x[w] + (half * dt) * k; acc[w] = acc[w] + (b2 * dt) * k
    k = _rhs(B[2:n - 2]);       w = slice(3, n - 3)
    A[w] = x[w] + (1.0 * dt) * k; acc[w] = acc[w] + (b2 * dt) * k
    k = _rhs(A[3:n - 3]);       w = slice(4, n - 4)
    x[w] = acc[w] + (b1 * dt) * k
    return x


def _worker(rank, world, port, n_total, steps, dt, x0, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ghost = 4
    b, e = shard_range(n_total, rank, world)
    slab = torch.zeros((1, (e - b) + 2 * ghost), dtype=torch.float64)
    slab[0, ghost:ghost + (e - b)] = torch.from_numpy(x0[b:e])
    # ghost cells after one exchange = neighbours' boundary cells on the periodic ring
    exchange_halos(slab, ghost, rank, world)
    want_l = x0[(np.arange(b - ghost, b)) % n_total]
    want_r = x0[(np.arange(e, e + ghost)) % n_total]
    ok_halo = np.array_equal(slab[0, :ghost].numpy(), want_l) and np.array_equal(slab[0, -ghost:].numpy(), want_r)
    for _ in range(steps):
        exchange_halos(slab, ghost, rank, world)
        slab[0] = torch.from_numpy(_sharded_rk4_step(slab[0].numpy().copy(), ghost, dt))
    q.put((rank, ok_halo, slab[0, ghost:ghost + (e - b)].numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _worker_interval(rank, world, port, n_total, steps, dt, x0, interval, q):
    """The scheme of odeintr_lattice_shard_run with the tiled kernels: ghost zone `interval` x 4 cells wide, one ring
    exchange per `interval` steps; every step consumes 4 ghost cells per side and recomputes the rest."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ghost = 4 * interval
    b, e = shard_range(n_total, rank, world)
    slab = torch.zeros((1, (e - b) + 2 * ghost), dtype=torch.float64)
    slab[0, ghost:ghost + (e - b)] = torch.from_numpy(x0[b:e])
    for s in range(steps):
        phase = s % interval
        if phase == 0:
            exchange_halos(slab, ghost, rank, world)
        # valid cells before the step: owned +- (ghost - 4 * phase); the step leaves owned +- (ghost - 4 * (phase + 1))
        lo = 4 * phase
        window = slab[0, lo:slab.shape[1] - lo].numpy().copy()
        slab[0, lo:slab.shape[1] - lo] = torch.from_numpy(_sharded_rk4_step(window, 4, dt))
    q.put((rank, True, slab[0, ghost:ghost + (e - b)].numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("interval", [3, 8])
def test_two_rank_exchange_every_k_steps_matches_oracle(interval):
    # communication-avoiding variant: bit-identical to the oracle whatever the exchange interval (13 steps: the last
    # round is incomplete)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import build_oracle as bo
    from systems import BISTABLE_1D, BISTABLE_PARS

    n_total, steps, dt = 211, 13, 0.1
    x0 = np.random.default_rng(23).uniform(0, 1, n_total)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_interval, args=(r, 2, port, n_total, steps, dt, x0, interval, q)) for r in range(2)]
    for p in procs:
        p.start()
    parts = dict()
    for _ in range(2):
        rank, _, part = q.get(timeout=240)
        parts[rank] = part
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.concatenate([parts[0], parts[1]])
    ora = bo.build(BISTABLE_1D, BISTABLE_PARS, True, "rk4", sys_dim=n_total)
    rec = ora.integrate_const(x0, 0.0, steps * dt + dt / 2, dt)
    assert rec["rows"] == steps + 1
    assert np.array_equal(got.view(np.uint64), rec["rec"][-1].view(np.uint64))


def _window_step(buf, lo, hi, dt):
    """cells [lo, hi) of the next state from buf (valid on [lo - 4, hi + 4)), as a new array of hi - lo values"""
    w = buf[lo - 4:hi + 4].copy()
    return _sharded_rk4_step(w, 4, dt)[4:-4]


def _worker_pairs(rank, world, port, n_total, steps, dt, x0, interval, q):
    """odeintr_lattice_shard_run with FUSED STEP PAIRS (abi.cu: slab_pair_possible / slab_tiled_pair), in the slab's
    local numbering: two consecutive steps of a round produce the middle [ilo, ihi) in one pass that reads `cur` only
    (a margin of 2 x 4 cells), the edges [lo1, ilo + 4) u [ihi - 4, hi1) go cur -> mid, then [lo2, ilo) u [ihi, hi2)
    mid -> other; steps a round cannot pair are single steps.  The region arithmetic is the C-ABI's."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m1, halo = 4, 1
    g = m1 * interval
    b, e = shard_range(n_total, rank, world)
    n = e - b
    cur = np.zeros(n + 2 * g)
    cur[g:g + n] = x0[b:e]
    c0 = g  # local index of the first owned cell
    e_lo, e_hi = c0 + g, c0 + n - g
    ilo, ihi = max(e_lo, c0 + 2 * m1 + halo), min(e_hi, c0 + n - 2 * m1 - halo)
    assert ihi - ilo > 4 * m1
    s = 0
    while s < steps:
        phase = s % interval
        if phase == 0:
            t = torch.from_numpy(cur).reshape(1, -1)
            exchange_halos(t, g, rank, world)
            cur = t.numpy().reshape(-1).copy()
        lo1, hi1 = c0 - g + m1 * (phase + 1), c0 + n + g - m1 * (phase + 1)
        other = np.full_like(cur, np.nan)
        if phase + 1 < interval and s + 1 < steps:
            lo2, hi2 = lo1 + m1, hi1 - m1
            mid = np.full_like(cur, np.nan)
            # edges, first step: everything the second step of the edges reads
            mid[lo1:min(ilo + m1, hi1)] = _window_step(cur, lo1, min(ilo + m1, hi1), dt)
            mid[max(ihi - m1, lo1):hi1] = _window_step(cur, max(ihi - m1, lo1), hi1, dt)
            other[lo2:ilo] = _window_step(mid, lo2, ilo, dt)
            other[ihi:hi2] = _window_step(mid, ihi, hi2, dt)
            # the middle: both steps from `cur` alone (what the fused pass keeps in registers)
            first = np.full_like(cur, np.nan)
            first[ilo - m1:ihi + m1] = _window_step(cur, ilo - m1, ihi + m1, dt)
            other[ilo:ihi] = _window_step(first, ilo, ihi, dt)
            s += 2
        else:
            other[lo1:hi1] = _window_step(cur, lo1, hi1, dt)
            s += 1
        cur = other
    q.put((rank, True, cur[c0:c0 + n].copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("interval,steps", [(8, 19), (3, 7), (2, 5)])
def test_two_rank_fused_step_pairs_match_oracle(interval, steps):
    # the pair schedule of odeintr_lattice_shard_run: whatever the round length (odd rounds leave a single step, calls
    # that end mid-round too), no launch ever reads a cell that is not valid (NaN would spread) and the owned cells carry
    # the oracle's bits
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import build_oracle as bo
    from systems import BISTABLE_1D, BISTABLE_PARS

    n_total, dt = 331, 0.1
    x0 = np.random.default_rng(29).uniform(0, 1, n_total)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_pairs, args=(r, 2, port, n_total, steps, dt, x0, interval, q)) for r in range(2)]
    for p in procs:
        p.start()
    parts = dict()
    for _ in range(2):
        rank, _, part = q.get(timeout=240)
        parts[rank] = part
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.concatenate([parts[0], parts[1]])
    ora = bo.build(BISTABLE_1D, BISTABLE_PARS, True, "rk4", sys_dim=n_total)
    rec = ora.integrate_const(x0, 0.0, steps * dt + dt / 2, dt)
    assert rec["rows"] == steps + 1
    assert np.array_equal(got.view(np.uint64), rec["rec"][-1].view(np.uint64))


@pytest.mark.timeout(300)
def test_two_rank_wide_halo_rk4_matches_oracle():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import build_oracle as bo
    from systems import BISTABLE_1D, BISTABLE_PARS

    n_total, steps, dt = 203, 12, 0.1
    rng = np.random.default_rng(21)
    x0 = rng.integers(0, 2, n_total).astype(np.float64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, steps, dt, x0, q)) for r in range(2)]
    for p in procs:
        p.start()
    parts = dict()
    for _ in range(2):
        rank, ok_halo, part = q.get(timeout=240)
        assert ok_halo
        parts[rank] = part
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.concatenate([parts[0], parts[1]])
    ora = bo.build(BISTABLE_1D, BISTABLE_PARS, True, "rk4", sys_dim=n_total)
    ref = ora.integrate_const(x0, 0.0, steps * dt, dt, record=False)
    # 12 steps of 0.1: integrate_const takes exactly 12 steps here
    assert np.array_equal(got.view(np.uint64), ref["x"].view(np.uint64))


def test_single_rank_exchange_wraps_onto_itself():
    x = torch.arange(10, dtype=torch.float64).reshape(1, 10).clone()  # ghost 2: owned cells are 2..7
    exchange_halos(x, 2, 0, 1)
    assert x[0].tolist() == [6.0, 7.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 2.0, 3.0]
