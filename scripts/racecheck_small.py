"""Small runs of every shared-memory kernel for `compute-sanitizer --tool racecheck` / `--tool memcheck`."""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import odeintr_b200
from systems import *
rng = np.random.default_rng(5)
env = {}
n = 1 << 14
odeintr_b200.compile_sys_openmp("t1", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n, env=env)
env["t1_no_record"](rng.uniform(0, 1, n), 0.3, 0.1); print("1-D rk4 tiled (TMA + general) ok", flush=True)
n = 2 * 6001
odeintr_b200.compile_sys_openmp("t2", CHAIN, CHAIN_PARS, True, method="rk4", sys_dim=n, globals=CHAIN_GLOBALS, env=env)
env["t2_no_record"](rng.uniform(-1, 1, n), 0.3, 0.1); print("1-D rk4 tiled, two fields, odd length (plain loads) ok", flush=True)
m = 160
odeintr_b200.compile_sys_openmp("t3", BISTABLE_2D, BISTABLE_PARS, True, method="rk4", sys_dim=m * m, env=env)
env["t3_no_record"](rng.uniform(0, 1, m * m), 0.2, 0.1); print("2-D rk4 tiled ok", flush=True)
n = 1 << 13
odeintr_b200.compile_sys_openmp("t4", BISTABLE_1D, BISTABLE_PARS, True, method="rk5_i", sys_dim=n, env=env)
env["t4"](rng.uniform(0, 1, n), 1.0, 0.25); print("dopri5 tiled ok", flush=True)
odeintr_b200.compile_sys_openmp("t5", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=300, env=env)
env["t5_no_record"](rng.uniform(0, 1, (300, 6)), 0.3, 0.1); print("ensemble, block per trajectory ok", flush=True)
odeintr_b200.compile_sys_openmp("t6", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=32, env=env)
env["t6_no_record"](rng.uniform(0, 1, (32, 10)), 0.3, 0.1); print("ensemble, warp per trajectory ok", flush=True)
