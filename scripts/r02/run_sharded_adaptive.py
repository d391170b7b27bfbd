"""torchrun entry: adaptive dopri5 (compile_sys_openmp's default method rk5_i) on a lattice slab-sharded over N GPUs:
one temporally blocked attempt per slab, the error norm all-reduced (ncclAllReduce max, 16 bytes), ghost cells of x and
of its derivative exchanged after every accepted step.  Checks the gathered record bit for bit against rank 0's
unsharded run on a 1 000 003-cell lattice (step counts included), then times bistable_at(0, 1, 5, 10) on a large one.
Usage: torchrun --nproc-per-node N scripts/r02/run_sharded_adaptive.py [log2_cells_per_gpu]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import odeintr_b200
from odeintr_b200 import _abi
from odeintr_b200.distributed import ShardedLattice, shard_range
from systems import BISTABLE_1D, BISTABLE_PARS

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _abi.load()
assert lib.odeintr_set_device(local) == 0
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 25

def init(f, cells):
    return 0.5 + 0.4 * np.sin(cells * 0.001) * np.cos(cells * 0.0137)

for method in ("rk5_i", "rk5_a"):
    n_small = 1_000_003
    kw = dict(method=method, sys_dim=n_small, atol=1e-7, rtol=1e-7)
    code = odeintr_b200.compile_sys_openmp("bs", BISTABLE_1D, BISTABLE_PARS, True, **kw)
    lat = ShardedLattice(code.system, n_small, rank=rank, world=world)
    lat.set_state_from(init)
    at = np.array([0.0, 0.5, 1.0 + 1e-3, 2.75])
    t, x = lat.integrate_times(at, 0.1)
    acc, rej = lat.stats()
    b, e = shard_range(n_small, rank, world)
    ok = torch.ones(1, dtype=torch.int32, device="cuda")
    env = {}
    code1 = odeintr_b200.compile_sys_openmp("bs1", BISTABLE_1D, BISTABLE_PARS, True, env=env, **kw)
    ref = env["bs1_at"](init(0, np.arange(n_small)), at, 0.1)   # every rank runs the unsharded reference on its own GPU
    st = code1.system.stats()
    mine = ref.iloc[:, 1:].to_numpy()[:, b:e]
    same = np.array_equal(x[:, 0, :].view(np.uint64), np.ascontiguousarray(mine).view(np.uint64)) and \
        acc == int(st["accepted"][0]) and rej == int(st["rejected"][0])
    ok[0] = int(same)
    if world > 1: dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded %s parity (world=%d, %d cells, accepted %d rejected %d): %s" %
              (method, world, n_small, acc, rej, "bit-exact" if int(ok.item()) else "MISMATCH"), flush=True)
    assert int(ok.item()) == 1
    code1.system.close(); del lat; code.system.close()

# throughput: 2^log2 cells per GPU, default method
n_total = (1 << log2) * world
code = odeintr_b200.compile_sys_openmp("bsl", BISTABLE_1D, BISTABLE_PARS, True, sys_dim=n_total)
lat = ShardedLattice(code.system, n_total, rank=rank, world=world)
at = np.array([0.0, 1.0, 5.0, 10.0])
for rep in range(2):
    lat.set_state_from(init)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    lat.integrate_times(at, 0.1) if False else lib.odeintr_integrate_times(code.system._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = torch.tensor([float(lib.odeintr_last_kernel_ms(code.system._h))], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
acc, rej = lat.stats()
if rank == 0:
    ms = float(ms.item())
    print(json.dumps({"metric": "lattice_dopri5_cell_attempts_per_s", "value": n_total * (acc + rej) / ms * 1e3, "n_gpus": world,
                      "cells_per_gpu": 1 << log2, "accepted": acc, "rejected": rej, "ms": ms, "ms_per_attempt": ms / (acc + rej),
                      "wall_ms": wall * 1e3, "launches": int(lib.odeintr_last_launches(code.system._h))}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
