"""Compiled system objects and the generated-function surface.

Mirrors what the reference's generated Rcpp module exports for a system called NAME
(``/root/reference/inst/templates/compile_sys_template.cpp:58-171`` and the ``pars_*_template.cpp``
files): ``NAME()``, ``NAME_adap()``, ``NAME_at()``, ``NAME_continue_at()``, ``NAME_no_record()``,
``NAME_reset_observer()``, ``NAME_get_state()``, ``NAME_set_state()``, ``NAME_get_output()``,
``NAME_set_params()``, ``NAME_get_params()`` -- with every numerical call going through the C-ABI
(``_abi.py`` -> ``libodeintr_b200.so``).  The only extension is that ``init`` / parameters may carry a
trailing instance axis: a state of shape ``(N, M)`` integrates an ensemble of M trajectories at once.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Dict, Optional

import numpy as np

from . import _abi
from .codegen import Generated, OdeintrError


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(_abi.c_dbl_p)


class EnsembleRecord:
    """Observer record of an ensemble run: ``time`` (rows,), ``x`` (rows, N, M)."""

    def __init__(self, time: np.ndarray, x: np.ndarray):
        self.time = time
        self.x = x

    @property
    def rows(self) -> int:
        return int(self.time.shape[0])

    def frame(self, instance: int = 0):
        """The data frame the reference returns for one trajectory: columns Time, X1..XN."""
        return _frame(self.time, self.x[:, :, instance])


def _frame(time: np.ndarray, cols: np.ndarray):
    import pandas as pd

    data = {"Time": np.asarray(time)}
    for v in range(cols.shape[1]):
        data["X%d" % (v + 1)] = np.ascontiguousarray(cols[:, v])
    return pd.DataFrame(data)


def proc_output(res):
    """Default output processor of the observer path, R/utils.R:242-262: ``res`` = [times, list of observer results]
    -> data frame with Time + the results' names (X1..Xk when unnamed); the raw list when the results are ragged."""
    import pandas as pd

    times, xs = res[0], res[1]
    if len(times) == 0:
        return None
    rows = [list(x.values()) if isinstance(x, dict) else list(np.atleast_1d(x)) for x in xs]
    if len({len(r) for r in rows}) != 1:
        return res
    first = xs[0]
    names = list(first.keys()) if isinstance(first, dict) else ["X%d" % (k + 1) for k in range(len(rows[0]))]
    data = {"Time": np.asarray(times)}
    for k, n in enumerate(names):
        data[n] = [r[k] for r in rows]
    return pd.DataFrame(data)


class CompiledSystem:
    """One NVRTC-compiled system resident on one GPU."""

    def __init__(self, gen: Generated, flags: int = 0, observer=None):
        # observer: the Python analogue of compile_sys(..., observer = f) (compile_sys_r_obs_template.cpp:44-53).
        # The reference calls the closure from inside the stepper loop; here the states are recorded on the GPU and
        # the closure is replayed over the stored samples on the host: same (x, t) sequence.
        self.observer = observer
        self.output_processor = proc_output
        self.gen = gen
        self.name = gen.name
        self.N = gen.sys_dim
        self.pars = gen.pars
        self.P = gen.pars.n_runtime
        self._lib = _abi.load()
        handle = C.c_void_p()
        if gen.stepper is None or gen.stepper.mode == "dense":
            flags |= _abi.FLAG_DENSE_OUTPUT
        if gen.family == "lattice" and gen.stepper.base == "euler":
            flags |= _abi.FLAG_LATTICE_EULER
        if gen.family == "lattice" and gen.stepper.method == "rk5":
            flags |= _abi.FLAG_LATTICE_RK5
        if gen.family == "lattice" and gen.stepper.base == "rk54":
            flags |= _abi.FLAG_LATTICE_CK54
        if gen.family == "lattice" and gen.stepper.base == "rk78":
            flags |= _abi.FLAG_LATTICE_RKF78
        if gen.family == "lattice" and gen.stepper.base in ("bs", "bsd"):
            flags |= _abi.FLAG_LATTICE_BS
        if gen.family == "lattice" and gen.stepper.mode != "plain":
            flags |= _abi.FLAG_LATTICE_ADAPTIVE
        rc = self._lib.odeintr_compile(gen.code.encode(), gen.name.encode(), self.N, self.P, gen.atol, gen.rtol,
                                       flags, C.byref(handle))
        if rc != 0:
            raise OdeintrError(_abi.last_error())
        self._h = handle
        if self.P:
            self._par_values = np.array(gen.pars.defaults, dtype=np.float64).reshape(self.P, 1)
            self._push_params()

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise OdeintrError(_abi.last_error())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.odeintr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _push_params(self):
        p = np.ascontiguousarray(self._par_values, dtype=np.float64)
        self._check(self._lib.odeintr_set_params(self._h, _dptr(p), p.size, p.shape[1]))

    @property
    def n_instances(self) -> int:
        return int(self._lib.odeintr_n_instances(self._h))

    # -- state (compile_sys_template.cpp:75-97) -----------------------------------------------------
    def set_state(self, new_state):
        x = np.asarray(new_state, dtype=np.float64)
        if x.ndim == 0:
            x = x.reshape(1)
        if x.ndim == 1:
            if x.shape[0] != self.N:
                raise OdeintrError("Invalid initial state")
            x = x.reshape(self.N, 1)
        if x.ndim != 2 or x.shape[0] != self.N or x.shape[1] == 0:
            raise OdeintrError("Invalid initial state")
        m = x.shape[1]
        # one parameter set per instance: a single-trajectory state is replicated over the sweep
        if self.P and self._par_values.shape[1] > 1 and m == 1:
            m = self._par_values.shape[1]
            x = np.repeat(x, m, axis=1)
        x = np.ascontiguousarray(x)
        self._check(self._lib.odeintr_set_state(self._h, _dptr(x), x.size, m))

    def get_state(self):
        m = self.n_instances
        x = np.empty((self.N, max(m, 1)), dtype=np.float64)
        self._check(self._lib.odeintr_get_state(self._h, _dptr(x), x.size))
        return x[:, 0].copy() if m == 1 else x

    def reset_observer(self):
        self._check(self._lib.odeintr_reset_observer(self._h))

    def get_output(self):
        if self.observer is not None:
            return self._get_observed_output()
        return self._get_state_output()

    def _get_observed_output(self):
        """NAME_get_output() of the observer template (compile_sys_r_obs_template.cpp:64-71): the observer sees every
        recorded (x, t); results of length 0 are skipped; the output processor shapes the list."""
        rec = self._get_state_output()
        if isinstance(rec, EnsembleRecord):
            raise OdeintrError("custom observers are replayed for single trajectories only")
        times, xs = [], []
        if rec is not None:
            t = rec["Time"].to_numpy()
            x = rec.iloc[:, 1:].to_numpy()
            for k in range(len(t)):
                r = self.observer(x[k].copy(), float(t[k]))
                if r is None or (hasattr(r, "__len__") and len(r) == 0):
                    continue
                xs.append(r)
                times.append(float(t[k]))
        return self.output_processor([times, xs])

    def _get_state_output(self):
        rows = int(self._lib.odeintr_output_rows(self._h))
        m = self.n_instances
        ragged = bool(self._lib.odeintr_output_is_ragged(self._h))
        if ragged:
            return self._get_ragged_output(rows, m)
        t = np.empty(rows, dtype=np.float64)
        self._check(self._lib.odeintr_get_output_times(self._h, _dptr(t), rows))
        if m <= 1:
            x = np.empty((self.N, rows), dtype=np.float64)
            if rows:
                self._check(self._lib.odeintr_get_output(self._h, _dptr(x), x.size, 0, 1))
            return _frame(t, x.T)
        x = np.empty((rows, self.N, m), dtype=np.float64)
        if rows:
            self._check(self._lib.odeintr_get_output_soa(self._h, _dptr(x), x.size))
        return EnsembleRecord(t, x)

    def _get_ragged_output(self, rows, m):
        """integrate_adaptive with an adaptive stepper: one row per accepted step, per instance."""
        counts = np.empty(m, dtype=np.int32)
        self._check(self._lib.odeintr_get_output_rows_per_instance(self._h, counts.ctypes.data_as(_abi.c_int_p), m))
        x = np.empty((rows, self.N, m), dtype=np.float64)
        self._check(self._lib.odeintr_get_output_soa(self._h, _dptr(x), x.size))
        times = np.empty((m, rows), dtype=np.float64)
        for i in range(m if m <= 64 else 0):
            self._check(self._lib.odeintr_get_output_times_per_instance(self._h, _dptr(times[i]), rows, i))
        if m == 1:
            n = int(counts[0])
            return _frame(times[0, :n], x[:n, :, 0])
        rec = EnsembleRecord(times[0] if m <= 64 else None, x)
        rec.rows_per_instance = counts
        rec.times_per_instance = times if m <= 64 else None
        return rec

    # -- parameters (pars_setter_template.cpp, pars_vec_setter_template.cpp) --------------------------
    def set_params(self, *args, **kw):
        if self.P == 0:
            raise OdeintrError("system has no run-time parameters")
        if self.pars.kind == "vector":
            if len(args) != 1 or kw:
                raise OdeintrError("Invalid parameter vector")
            p = np.asarray(args[0], dtype=np.float64)
            if p.ndim == 1:
                p = p.reshape(-1, 1)
            if p.ndim != 2 or p.shape[0] != self.P:
                raise OdeintrError("Invalid parameter vector")
            self._par_values = np.ascontiguousarray(p)
        else:
            names = self.pars.names
            given: Dict[str, object] = dict(zip(names, args))
            for k, v in kw.items():
                if k not in names:
                    raise OdeintrError("unused argument (%s)" % k)
                given[k] = v
            # R default arguments: parameters not named in the call are reset to their declared defaults
            # (pars_setter_template.cpp: `double a = dflt`); character pars have no default and are required
            cols = []
            m = 1
            for k, n in enumerate(names):
                if n in given:
                    v = np.asarray(given[n], dtype=np.float64).reshape(-1)
                elif self.pars.has_defaults:
                    v = np.array([self.pars.defaults[k]])
                else:
                    raise OdeintrError('argument "%s" is missing, with no default' % n)
                m = max(m, v.shape[0])
                cols.append(v)
            out = np.empty((self.P, m), dtype=np.float64)
            for k, v in enumerate(cols):
                if v.shape[0] not in (1, m):
                    raise OdeintrError("Invalid parameter vector")
                out[k, :] = v
            self._par_values = out
        self._push_params()

    def get_params(self):
        p = self._par_values
        if self.pars.kind == "vector":
            return p[:, 0].copy() if p.shape[1] == 1 else p.copy()
        return {n: (float(p[k, 0]) if p.shape[1] == 1 else p[k].copy()) for k, n in enumerate(self.pars.names)}

    # -- integration (compile_sys_template.cpp:99-171) ------------------------------------------------
    def run(self, init, duration, step_size=1.0, start=0.0, obs_stride=1):
        """NAME(init, duration, step_size, start): integrate_const with the observer."""
        if duration / step_size < 3:
            step_size = duration / 3
            warnings.warn("Step size too large -- adjusting")
        self.set_state(init)
        self.reset_observer()
        self._check(self._lib.odeintr_integrate_const(self._h, start, start + duration, step_size, obs_stride, 1))
        return self.get_output()

    def adap(self, init, duration, step_size=1.0, start=0.0):
        self.set_state(init)
        self.reset_observer()
        self._check(self._lib.odeintr_integrate_adaptive(self._h, start, start + duration, step_size, 1))
        return self.get_output()

    def _at(self, times, step_size, start):
        times = np.ascontiguousarray(np.asarray(times, dtype=np.float64).reshape(-1))
        if times.size == 0:
            raise OdeintrError("Invalid time specification")
        self._check(self._lib.odeintr_integrate_const(self._h, start, float(times[0]), step_size, 1, 0))
        self._check(self._lib.odeintr_integrate_times(self._h, _dptr(times), times.size, step_size))
        return self.get_output()

    def at(self, init, times, step_size=1.0, start=0.0):
        self.set_state(init)
        self.reset_observer()
        return self._at(times, step_size, start)

    def continue_at(self, times, step_size=1.0):
        rows = int(self._lib.odeintr_output_rows(self._h))
        if rows == 0:
            # the reference reads rec_t.back() of an empty vector here (undefined behaviour)
            raise OdeintrError("no previous record to continue from")
        t = np.empty(rows, dtype=np.float64)
        if self._lib.odeintr_output_is_ragged(self._h):
            # NAME_adap() with a controlled / dense stepper left one row per accepted step: the reference reads
            # rec_t.back() (compile_sys_template.cpp:130), the time of the trajectory's last row.  Every instance of an
            # ensemble ends at the same time, start + duration, so instance 0 answers for all.
            m = self.n_instances
            counts = np.empty(m, dtype=np.int32)
            self._check(self._lib.odeintr_get_output_rows_per_instance(self._h, counts.ctypes.data_as(_abi.c_int_p), m))
            self._check(self._lib.odeintr_get_output_times_per_instance(self._h, _dptr(t), rows, 0))
            t = t[:int(counts[0])]
        else:
            self._check(self._lib.odeintr_get_output_times(self._h, _dptr(t), rows))
        start = float(t[-1])
        self.reset_observer()
        return self._at(times, step_size, start)

    def no_record(self, init, duration, step_size=1.0, start=0.0):
        self.set_state(init)
        self._check(self._lib.odeintr_integrate_adaptive(self._h, start, start + duration, step_size, 0))
        return self.get_state()

    def run_host(self, init, duration, step_size=1.0, start=0.0, obs_stride=1, out=None, x_final=None, pars=None,
                 chunk=0, record=True):
        """NAME() for ensembles that live in host memory (and may exceed device memory): the instances are
        streamed through the GPU in chunks with copies and kernels overlapped (odeintr_integrate_const_host).
        ``init`` is (N, M); ``out`` / ``x_final`` may be preallocated (pinned) arrays.  Returns
        (EnsembleRecord or None, x_final)."""
        x = np.ascontiguousarray(np.asarray(init, dtype=np.float64))
        if x.ndim != 2 or x.shape[0] != self.N or x.shape[1] == 0:
            raise OdeintrError("Invalid initial state")
        m = x.shape[1]
        rows = int(self._lib.odeintr_const_rows(start, start + duration, step_size, obs_stride)) if record else 0
        if record and out is None:
            out = np.empty((rows, self.N, m), dtype=np.float64)
        if x_final is None:
            x_final = np.empty((self.N, m), dtype=np.float64)
        p = None
        par_sets = 0
        if self.P:
            p = np.ascontiguousarray(self._par_values if pars is None else np.asarray(pars, dtype=np.float64))
            par_sets = p.shape[1]
        self._check(self._lib.odeintr_integrate_const_host(
            self._h, _dptr(x), m, None if p is None else _dptr(p), par_sets, start, start + duration, step_size,
            obs_stride, None if not record else _dptr(out), out.size if record else 0, _dptr(x_final), chunk))
        if not record:
            return None, x_final
        t = np.array([start + (s * obs_stride) * step_size for s in range(rows - 1)] +
                     [start + self.last_steps() * step_size])
        return EnsembleRecord(t, out), x_final

    def set_instance_order(self, order="index"):
        """How the instances of an ensemble are dealt to the lanes by the adaptive steppers (odeintr_set_instance_order):
        "index" (default), "cost" (sorted by a pilot's cost proxy before every integrate call: for shuffled / random
        ensembles) or an explicit permutation.  Results do not depend on it, only the time."""
        if isinstance(order, str):
            mode = {"index": _abi.ORDER_INDEX, "cost": _abi.ORDER_COST_PROXY}.get(order)
            if mode is None:
                raise OdeintrError("Invalid instance order")
            self._check(self._lib.odeintr_set_instance_order(self._h, mode, None, 0))
            return
        perm = np.ascontiguousarray(np.asarray(order, dtype=np.int32).reshape(-1))
        self._check(self._lib.odeintr_set_instance_order(self._h, _abi.ORDER_PERMUTATION,
                                                         perm.ctypes.data_as(_abi.c_int_p), perm.size))

    def get_instance_order(self):
        perm = np.empty(self.n_instances, dtype=np.int32)
        self._check(self._lib.odeintr_get_instance_order(self._h, perm.ctypes.data_as(_abi.c_int_p), perm.size))
        return perm

    def set_halo(self, halo: int, fields: int = 1):
        """Stencil radius of a large system for the tiled rk4 kernels: declared (> 0), measured (< 0) or
        none (0 = stage-per-launch kernels) -- odeintr_lattice_set_halo."""
        self._check(self._lib.odeintr_lattice_set_halo(self._h, halo, fields))

    def get_halo(self) -> int:
        """Radius the tiled rk4 kernels run with, -1 when the stage-per-launch kernels are in use."""
        return int(self._lib.odeintr_lattice_get_halo(self._h))

    # -- measurement ----------------------------------------------------------------------------------
    def last_kernel_ms(self) -> float:
        return float(self._lib.odeintr_last_kernel_ms(self._h))

    def last_steps(self) -> int:
        return int(self._lib.odeintr_last_steps(self._h))

    def last_launches(self) -> int:
        return int(self._lib.odeintr_last_launches(self._h))

    def kernel_attributes(self, kernel: str):
        r, l, s, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.odeintr_kernel_attributes(self._h, kernel.encode(), C.byref(r), C.byref(l), C.byref(s),
                                                        C.byref(t)))
        return {"registers": r.value, "local_bytes": l.value, "shared_bytes": s.value, "max_threads": t.value}

    def stats(self):
        m = self.n_instances
        acc = np.empty(m, dtype=np.int64)
        rej = np.empty(m, dtype=np.int64)
        st = np.empty(m, dtype=np.int32)
        self._check(self._lib.odeintr_get_stats(self._h, acc.ctypes.data_as(_abi.c_ll_p), rej.ctypes.data_as(_abi.c_ll_p),
                                                st.ctypes.data_as(_abi.c_int_p), m))
        return {"accepted": acc, "rejected": rej, "status": st}

    # -- the generated-function table ---------------------------------------------------------------------
    def exports(self) -> Dict[str, object]:
        n = self.name
        fns = {
            n: self.run,
            n + "_adap": self.adap,
            n + "_at": self.at,
            n + "_continue_at": self.continue_at,
            n + "_no_record": self.no_record,
            n + "_reset_observer": self.reset_observer,
            n + "_get_state": self.get_state,
            n + "_set_state": self.set_state,
            n + "_get_output": self.get_output,
        }
        if self.P:
            fns[n + "_set_params"] = self.set_params
            fns[n + "_get_params"] = self.get_params
        if self.observer is not None:  # compile_sys_r_obs_template.cpp:171-197
            fns[n + "_get_observer"] = lambda: self.observer
            fns[n + "_set_observer"] = lambda f: setattr(self, "observer", f)
            fns[n + "_get_output_processor"] = lambda: self.output_processor
            fns[n + "_set_output_processor"] = lambda f: setattr(self, "output_processor", f)
        return fns
