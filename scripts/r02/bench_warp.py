#!/usr/bin/env python3
"""K4w (warp-autonomous tiled rk4, lattice_warp.cuh) against K4t (block-tiled TMA kernel) and the staged schedule on
BASELINE.json configs[4]-sized lattices: 2^28 states, bistable (1 field) and oscillator chain (2 fields).
Final states are compared bit for bit between the paths.  Usage: python scripts/r02/bench_warp.py [steps]"""
import ctypes as C, hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import odeintr_b200
from odeintr_b200 import _abi
from systems import BISTABLE_1D, BISTABLE_PARS, CHAIN, CHAIN_GLOBALS, CHAIN_PARS
lib = _abi.load()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dp = C.c_double(); ms = C.c_double()
lib.odeintr_fp64_probe(1, 0, C.byref(dp), C.byref(ms))
n = 1 << 28
cases = {
    "bistable": (BISTABLE_1D, BISTABLE_PARS, "", 50.0, (np.arange(n) % 7 < 3).astype(np.float64)),
    "chain": (CHAIN, CHAIN_PARS, CHAIN_GLOBALS, 32.0, None),
}
x_chain = np.zeros(n); x_chain[:n // 2] = np.sin(2 * np.pi * 8 * np.arange(n // 2) / (n // 2))
for name in (sys.argv[2:] or ["bistable", "chain"]):
    src, pars, glob, dpi, x0 = cases[name]
    if x0 is None: x0 = x_chain
    s = odeintr_b200.compile_sys_openmp(name, src, pars, True, method="rk4", sys_dim=n, globals=glob).system
    digests = {}
    for label, env in (("warp_tma", {"ODEINTR_FUSE": "0"}), ("warp_tma_two_steps_per_pass", {"ODEINTR_FUSE": "1"}),
                       ("warp_plain", {"ODEINTR_NO_TMA": "1", "ODEINTR_FUSE": "0"}), ("block_tma", {"ODEINTR_NO_WARP": "1"})):
        for k in ("ODEINTR_NO_TMA", "ODEINTR_NO_WARP", "ODEINTR_FUSE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        best = None
        for rep in range(3):
            s.set_state(x0)
            assert lib.odeintr_integrate_const(s._h, 0.0, steps * 0.01, 0.01, 1, 0) == 0, _abi.last_error()
            best = s.last_kernel_ms() if best is None else min(best, s.last_kernel_ms())
        fin = s.get_state()
        digests[label] = hashlib.sha1(fin.tobytes()).hexdigest()
        rate = n * steps / best * 1e3
        print(json.dumps({"case": name, "path": label, "ms_per_step": best / steps, "rate": rate, "fp64_frac": rate * dpi / dp.value,
                          "hbm_GBps": rate * (8 if "two_steps" in label else 16) / 1e9, "launches": s.last_launches(), "sha1": digests[label][:12]}), flush=True)
    print(name, "all paths bit-identical:", len(set(digests.values())) == 1, flush=True)
    for k in ("odeintr_lattice_rk4_warp_h1", "odeintr_lattice_rk4_tiled_tma"):
        print(k, s.kernel_attributes(k), flush=True)
    s.close()
