"""Scratch: register cap / block size sweep for the adaptive ensemble kernels (configs 3 and 4), 10 M instances."""
import ctypes as C, os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import odeintr_b200
    from odeintr_b200 import _abi
    from systems import VDP, ROBERTSON, ROBERTSON_PARS
    lib = _abi.load()
    m = 10_000_000
    which = sys.argv[2]
    if which == "vdp":
        s = odeintr_b200.compile_sys("vpol", VDP, "mu", method="rk5_i", atol=1e-8, rtol=1e-8).system
        lo = (C.c_double * 1)(0.5); hi = (C.c_double * 1)(2.0)
        lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi)
        times = np.arange(0, 20.0001, 0.2); X = np.repeat(np.array([2.0, 0.0])[:, None], m, axis=1); dt = 0.01
    else:
        s = odeintr_b200.compile_implicit("rob", ROBERTSON, ROBERTSON_PARS).system
        lo = (C.c_double * 3)(0.02, 1e4, 3e7); hi = (C.c_double * 3)(0.08, 1e4, 3e7)
        lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi)
        times = 10 ** np.linspace(-5, 5, 41); X = np.repeat(np.array([1.0, 0.0, 0.0])[:, None], m, axis=1); dt = 1e-6
    best = 1e30
    for rep in range(3):
        s.set_state(X)
        assert lib.odeintr_integrate_times(s._h, times.ctypes.data_as(_abi.c_dbl_p), times.size, dt) == 0, _abi.last_error()
        best = min(best, s.last_kernel_ms())
    a = s.kernel_attributes("odeintr_adaptive_ensemble")
    st = s.stats()
    print(json.dumps({"which": which, "regs": a["registers"], "local": a["local_bytes"], "ms": best, "acc": int(st["accepted"].sum()),
                      "minblocks": os.environ.get("ODEINTR_MIN_BLOCKS"), "block": os.environ.get("ODEINTR_BLOCK")}), flush=True)
    sys.exit(0)
for which in ("vdp", "rob"):
    for maxr, block in ((None, None), ("5", None), ("6", None), ("8", None), ("10", None), ("2", "256"), ("3", "256"), ("4", "256"), ("12", "64"), ("16", "64")):
        env = dict(os.environ)
        if maxr: env["ODEINTR_MIN_BLOCKS"] = maxr
        if block: env["ODEINTR_BLOCK"] = block
        r = subprocess.run([sys.executable, __file__, "child", which], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("FAIL " + r.stderr[-300:]), flush=True)
