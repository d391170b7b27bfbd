import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import odeintr_b200
from oracle import build_oracle as bo
from systems import *
n = 2053
rng = np.random.default_rng(51); x0 = rng.uniform(0, 1, n)
for method in ("rk5_a", "rk5_i"):
    env = {}
    code = odeintr_b200.compile_sys_openmp("bis", BISTABLE_1D, BISTABLE_PARS, True, sys_dim=n, env=env, method=method)
    ora = bo.build(BISTABLE_1D, BISTABLE_PARS, True, method, sys_dim=n)
    for T, dt in ((0.01, 0.01), (0.5, 0.1), (4.2, 0.3)):
        fin = env["bis_no_record"](x0, T, dt)
        ref = ora.integrate_adaptive(x0, 0, T, dt)
        st = code.system.stats()
        print(method, "no_record", T, dt, "maxerr", np.max(np.abs(fin - ref["x"])), "acc", st["accepted"][0], ref["accepted"], "rej", st["rejected"][0], ref["rejected"], flush=True)
    df = env["bis"](x0, 6.0, 0.5)
    ref = ora.integrate_const(x0, 0, 6.0, 0.5)
    got = df.iloc[:, 1:].to_numpy()
    st = code.system.stats()
    print(method, "const rows", len(df), ref["rows"], "per-row maxerr", [float("%.2g" % np.max(np.abs(got[i] - ref["rec"][i]))) for i in range(len(df))], "acc", st["accepted"][0], ref["accepted"], flush=True)
