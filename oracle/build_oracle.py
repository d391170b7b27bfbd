"""oracle/build_oracle.py -- TEST INFRASTRUCTURE ONLY.

Builds a CPU oracle for a user system: fills oracle/oracle_template.cpp with the same snippet the
product compiles for the GPU and compiles it with ``g++ -O2 -ffp-contract=off`` against
oracle/odeint_restated.hpp (the restatement of Boost odeint).  This is what the
reference does with Rcpp::sourceCpp + Boost (/root/reference/R/odeintr.R:340), minus R and Boost, which
are not available in this image.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module; the product (odeintr_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import pathlib
import subprocess
import sys

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_BUILD = _HERE / "_build"
sys.path.insert(0, str(_HERE.parent))
from odeintr_b200 import codegen  # noqa: E402  (text helpers only: number formatting, sys_dim inference)

METHOD_IDS = {"euler": 0, "rk4": 1, "rk54": 2, "rk5": 3, "rk78": 4, "rk78_a": 13, "bs": 14, "bsd": 15, "rk5_a": 10, "rk5_i": 11, "rk54_a": 12, "rosenbrock4": 20}

c_dbl_p = C.POINTER(C.c_double)
c_long_p = C.POINTER(C.c_long)


def _dp(a):
    return a.ctypes.data_as(c_dbl_p) if a is not None else None


def _pars_text(pars, const, shared=False):
    """shared=True: plain statics (a large-system snippet runs its own `#pragma omp parallel for`, whose worker
    threads must see the values); otherwise thread_local so the ensemble loop can set them per instance."""
    p = codegen.handle_pars(pars, const)
    if p.kind == "none":
        return "", "", 0
    if p.kind == "vector":
        decl = "static %sstd::array<double, %d> pars;" % ("" if shared else "thread_local ", len(p.names))
    elif p.const:
        decl = "static " + p.globals
    else:
        decl = "static %s" % ("" if shared else "thread_local ") + p.globals
    return decl, p.load.replace("\n      ", "\n  "), p.n_runtime


def generate(sys_src, pars=None, const=False, method="rk4", sys_dim=-1, atol=1e-6, rtol=1e-6, globals="",
             jacobian=None) -> str:
    sys_src = sys_src if isinstance(sys_src, str) else "; ".join(sys_src)
    if sys_dim is None or sys_dim < 1:
        sys_dim = codegen.get_sys_dim(sys_src)
    if sys_dim is None:
        sys_src = codegen.vectorize_1d_sys(sys_src)
        sys_dim = 1
    if method == "rosenbrock4" and jacobian is None:
        jacobian = codegen.JacobianCpp(sys_src, sys_dim)
    decl, setp, npars = _pars_text(pars, const, shared="#pragma omp" in sys_src)
    text = (_HERE / "oracle_template.cpp").read_text()
    subs = {
        "@SYS_SIZE@": str(int(sys_dim)),
        "@METHOD@": str(METHOD_IDS[method]),
        "@ATOL@": repr(float(atol)),
        "@RTOL@": repr(float(rtol)),
        "@NPARS@": str(npars),
        "@GLOBALS@": decl + "\n" + globals,
        "@SET_PARS@": setp,
        "@JACOBIAN@": jacobian or "",
        "@SYS@": sys_src,
    }
    # the header comment names the placeholders; substitute only below it
    head, body = text.split("#include <cmath>", 1)
    body = "#include <cmath>" + body
    for k, v in subs.items():
        body = body.replace(k, v)
    return head + body


class Oracle:
    def __init__(self, lib: C.CDLL, path: pathlib.Path):
        self.lib = lib
        self.path = path
        lib.oracle_sys_dim.restype = C.c_int
        lib.oracle_n_pars.restype = C.c_int
        lib.oracle_max_threads.restype = C.c_int
        lib.oracle_run.restype = C.c_long
        lib.oracle_run.argtypes = [C.c_int, c_dbl_p, c_dbl_p, C.c_double, C.c_double, C.c_double, c_dbl_p, C.c_size_t,
                                   c_dbl_p, c_dbl_p, C.c_size_t, c_long_p]
        lib.oracle_ensemble.restype = C.c_long
        lib.oracle_ensemble.argtypes = [C.c_int, c_dbl_p, c_dbl_p, C.c_size_t, C.c_int, C.c_double, C.c_double,
                                        C.c_double, c_dbl_p, C.c_size_t, C.c_long, c_dbl_p, C.c_size_t, c_long_p,
                                        c_long_p, C.c_int]
        lib.oracle_rosenbrock4_step.restype = C.c_int
        lib.oracle_rosenbrock4_step.argtypes = [c_dbl_p, c_dbl_p, C.c_double, C.c_double, C.c_double, c_dbl_p, c_dbl_p,
                                                c_dbl_p]
        self.N = lib.oracle_sys_dim()
        self.P = lib.oracle_n_pars()

    def max_threads(self) -> int:
        return int(self.lib.oracle_max_threads())

    def _run(self, kind, x0, pars, t0, t1, dt, times, cap, record=True):
        x = np.array(x0, dtype=np.float64).reshape(self.N).copy()
        p = None if pars is None else np.ascontiguousarray(np.asarray(pars, dtype=np.float64).reshape(-1))
        tt = None if times is None else np.ascontiguousarray(np.asarray(times, dtype=np.float64))
        rec_t = np.empty(cap, dtype=np.float64) if record else None
        rec_x = np.empty((cap, self.N), dtype=np.float64) if record else None
        cnt = np.zeros(3, dtype=np.int64)
        rows = self.lib.oracle_run(kind, _dp(x), _dp(p), t0, t1, dt, _dp(tt), 0 if tt is None else tt.size,
                                   _dp(rec_t), _dp(rec_x), cap if record else 0, cnt.ctypes.data_as(c_long_p))
        if rows < 0:
            raise RuntimeError("Max number of iterations exceeded (500). A new step size was not found.")
        if record and rows > cap:
            return self._run(kind, x0, pars, t0, t1, dt, times, int(rows), record)
        out = {"x": x, "accepted": int(cnt[0]), "rejected": int(cnt[1]), "rhs": int(cnt[2]), "rows": int(rows)}
        if record:
            out["t"] = rec_t[:rows].copy()
            out["rec"] = rec_x[:rows].copy()
        return out

    def rosenbrock4_step(self, x, t, dt, theta=0.5, pars=None):
        """One raw rosenbrock4 step with a fixed step size: (x_new, x_err, dense(t + theta dt))."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(self.N))
        p = None if pars is None else np.ascontiguousarray(np.asarray(pars, dtype=np.float64).reshape(-1))
        xo, xe, xd = (np.empty(self.N) for _ in range(3))
        if self.lib.oracle_rosenbrock4_step(_dp(x), _dp(p), t, dt, theta, _dp(xo), _dp(xe), _dp(xd)) != 0:
            raise RuntimeError("oracle was not built with rosenbrock4")
        return xo, xe, xd

    # single trajectory -- name(), name_at(), name_adap()/name_no_record() of the reference
    def integrate_const(self, x0, t0, t1, dt, pars=None, record=True):
        cap = int(abs((t1 - t0) / dt)) + 8 if record else 0
        return self._run(0, x0, pars, t0, t1, dt, None, cap, record)

    def integrate_times(self, x0, times, dt, pars=None):
        return self._run(1, x0, pars, 0.0, 0.0, dt, times, len(times) + 1, True)

    def integrate_adaptive(self, x0, t0, t1, dt, pars=None, record=False):
        cap = int(abs((t1 - t0) / dt)) + 1024 if record else 0
        return self._run(2, x0, pars, t0, t1, dt, None, cap, record)

    def ensemble(self, kind, X, P=None, t0=0.0, t1=0.0, dt=1.0, times=None, obs_stride=1, out_rows=0, threads=1):
        """X: (N, m) array, advanced in place; returns dict(x=..., out=(rows,N,m) or None, accepted, rejected)."""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        m = X.shape[1]
        par_inc = 0
        p = None
        if P is not None:
            p = np.ascontiguousarray(np.asarray(P, dtype=np.float64))
            par_inc = 1 if (p.ndim == 2 and p.shape[1] == m and m > 1) else 0
        tt = None if times is None else np.ascontiguousarray(np.asarray(times, dtype=np.float64))
        out = np.zeros((out_rows, self.N, m), dtype=np.float64) if out_rows else None
        acc = np.zeros(m, dtype=np.int64)
        rej = np.zeros(m, dtype=np.int64)
        rows = self.lib.oracle_ensemble(kind, _dp(X), _dp(p), m, par_inc, t0, t1, dt, _dp(tt),
                                        0 if tt is None else tt.size, obs_stride, _dp(out), out_rows,
                                        acc.ctypes.data_as(c_long_p), rej.ctypes.data_as(c_long_p), threads)
        if rows < 0:
            raise RuntimeError("%d instances hit the 500-failed-steps overflow" % (-rows))
        return {"x": X, "out": None if out is None else out[:rows], "accepted": acc, "rejected": rej, "rows": rows}


def build(sys_src, pars=None, const=False, method="rk4", sys_dim=-1, atol=1e-6, rtol=1e-6, globals="",
          jacobian=None, openmp=True) -> Oracle:
    code = generate(sys_src, pars, const, method, sys_dim, atol, rtol, globals, jacobian)
    hdr = (_HERE / "odeint_restated.hpp").read_bytes()
    key = hashlib.sha1(code.encode() + hdr + (b"omp" if openmp else b"")).hexdigest()[:16]
    _BUILD.mkdir(exist_ok=True)
    so = _BUILD / ("oracle_%s.so" % key)
    if not so.exists():
        src = _BUILD / ("oracle_%s.cpp" % key)
        src.write_text(code)
        tmp = _BUILD / ("oracle_%s.%d.tmp.so" % (key, os.getpid()))
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", str(_HERE), str(src), "-o",
               str(tmp)]
        if openmp:
            cmd.insert(1, "-fopenmp")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stderr)
        os.replace(tmp, so)
    return Oracle(C.CDLL(str(so)), so)
