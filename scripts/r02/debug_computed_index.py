import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import odeintr_b200
from oracle import build_oracle as bo
n = 6 * 1024 + 77
srcs = {"computed": "for (int i = 0; i < N; ++i) dxdt[i] = D * (x[(i * 1 + N - 1) % N] + x[(i + 1) % N] - 2.0 * x[i * 1]);",
        "only_centre": "for (int i = 0; i < N; ++i) dxdt[i] = D * (x[(i + N - 1) % N] + x[(i + 1) % N] - 2.0 * x[i * 1]);",
        "only_left": "for (int i = 0; i < N; ++i) dxdt[i] = D * (x[(i * 1 + N - 1) % N] + x[(i + 1) % N] - 2.0 * x[i]);",
        "signed": "for (int i = 0; i < N; ++i) dxdt[i] = D * (x[(i * 1 + (int)N - 1) % (int)N] + x[(i + 1) % N] - 2.0 * x[i]);"}
x0 = np.random.default_rng(93).uniform(0, 1, n)
for name, src in srcs.items():
    ora = bo.build(src, {"D": 0.2}, True, "rk4", sys_dim=n)
    ref = ora.integrate_adaptive(x0, 0, 1.25, 0.1)["x"]
    for env in ({}, {"ODEINTR_NO_TMA": "1"}, {"ODEINTR_NO_WARP": "1"}, {"ODEINTR_TILED_GENERAL_ONLY": "1"}):
        for k in ("ODEINTR_NO_TMA", "ODEINTR_NO_WARP", "ODEINTR_TILED_GENERAL_ONLY"):
            os.environ.pop(k, None)
        os.environ.update(env)
        e = {}
        odeintr_b200.compile_sys_openmp("t", src, {"D": 0.2}, True, method="rk4", sys_dim=n, env=e, halo=1)
        try:
            fin = e["t_no_record"](x0, 1.25, 0.1)
            print(name, env, "bit-exact" if np.array_equal(fin.view(np.uint64), ref.view(np.uint64)) else "MISMATCH", flush=True)
        except Exception as ex:
            print(name, env, "ERROR", str(ex)[:90], flush=True)
