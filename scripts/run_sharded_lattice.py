"""torchrun entry: slab-sharded lattice (BASELINE.json configs[4]) over N GPUs: temporally blocked rk4 kernels on
every slab, NCCL ring halo exchange every SPE steps (env SPE, default 8; TILED=0 selects the stage-per-launch
kernels with one exchange per step).  Checks the gathered state bit for bit against rank 0's unsharded run on a small lattice, then times a
large one.  Usage: torchrun --nproc-per-node N scripts/run_sharded_lattice.py [log2_cells_per_gpu] [steps]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.distributed as dist
import odeintr_b200
from odeintr_b200 import _abi
from odeintr_b200.distributed import ShardedLattice
from systems import BISTABLE_1D, BISTABLE_PARS

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _abi.load()
assert lib.odeintr_set_device(local) == 0
if os.environ.get("TILED", "1") == "0":
    os.environ["ODEINTR_NO_TILED"] = "1"
spe = int(os.environ.get("SPE", "8"))
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 25
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50

def init(f, cells):
    return ((cells % 7) < 3).astype(np.float64)

# 1) parity on a small lattice
n_small = 1_000_003
code = odeintr_b200.compile_sys_openmp("bs", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n_small)
lat = ShardedLattice(code.system, n_small, rank=rank, world=world, steps_per_exchange=spe)
lat.set_state_from(init)
lat.step(10, 0.1, t0=0.0)
full = lat.gather_state().cpu().numpy().reshape(-1)
if rank == 0:
    code1 = odeintr_b200.compile_sys_openmp("bs1", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n_small)
    ref = code1.system.no_record(init(0, np.arange(n_small)), 1.0, 0.1)
    ok = bool(np.array_equal(full.view(np.uint64), ref.view(np.uint64)))
    print("sharded parity (world=%d, %d cells, 10 steps): %s" % (world, n_small, "bit-exact" if ok else "MISMATCH"), flush=True)
    assert ok

# 2) throughput, weak scaling: 2^log2 cells per GPU
n_total = (1 << log2) * world
code = odeintr_b200.compile_sys_openmp("bsl", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n_total)
lat = ShardedLattice(code.system, n_total, rank=rank, world=world, steps_per_exchange=spe)
lat.set_state_from(init)
lat.step(5, 0.01, t0=0.0)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(lat.stream)
lat.step(steps, 0.01)
e1.record(lat.stream); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(ms.item())
    print(json.dumps({"metric": "lattice_rk4_cell_steps_per_s", "value": n_total * steps / ms * 1e3, "n_gpus": world,
                      "cells_per_gpu": 1 << log2, "steps": steps, "ms_per_step": ms / steps,
                      "staged_schedule_GBps_per_gpu": 128.0 * (1 << log2) * steps / ms / 1e6,
                      "ghost_cells_per_side": lat.ghost, "launches": int(lib.odeintr_last_launches(code.system._h)),
                      "tiled": os.environ.get("ODEINTR_NO_TILED") is None, "steps_per_exchange": spe}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
