// odeintr_b200/csrc/kernels/integrate_device.cuh
// Per-trajectory use of the integrate loops (integrate_loops.cuh): every thread runs the loop of its own
// trajectory, so step sizes, rejected tries and the number of real steps between two observations are per
// instance.  This header adds the device observer.
#pragma once
#include "odeintr_b200/adaptive_steppers.cuh"
#include "odeintr_b200/rk78.cuh"
#include "odeintr_b200/bulirsch_stoer.cuh"
#include "odeintr_b200/bulirsch_stoer_dense.cuh"
#include "odeintr_b200/integrate_loops.cuh"

namespace odeintr_b200 {

// Observer: one SoA row per variable per sample, instance index fastest (coalesced across a warp whenever
// neighbouring trajectories reach the same sample together, which they do for smooth sweeps).
template <int N>
struct sample_writer {
  double* out;        // points at this instance's column of sample 0 (or null: null_observer)
  double* t_rec;      // per-instance sample times (ragged adaptive records) or null
  long long ld;
  int rows, max_rows;
  ODEINTR_DEV void operator()(const vec<N>& x, const double t) {
    if (out && rows < max_rows) {
#pragma unroll
      for (int v = 0; v < N; ++v) store_stream(out + ((long long)rows * N + v) * ld, x[v]);
      if (t_rec) store_stream(t_rec + (long long)rows * ld, t);
    }
    ++rows;
  }
};

// ---- warp-synchronous forms of the dense-output loops ------------------------------------------------------
// Same control flow as integrate_times / integrate_const / integrate_adaptive for dense-output steppers
// (integrate_loops.cuh), restructured so that ONE iteration of the loop makes at most one
// step attempt.  In Boost's form a rejected attempt is retried inside do_step; with 32 trajectories per warp
// nearly every do_step then costs the warp two attempts as soon as neighbouring instances reject independently
// (measured: a shuffled Van der Pol sweep ran 2.7x slower than the sorted one).  Here a rejecting lane simply
// retries in the next iteration, side by side with lanes that start their next step; the cheap bookkeeping
// (sample emission, branch decisions) is the only divergent part.  The sequence of operations of every
// individual trajectory is unchanged.
#define ODEINTR_NO_PROGRESS_LIMIT 8

template <class St, class Sys, int N, class Obs>
ODEINTR_DEV bool flat_times_dense(St& st, Sys& sys, vec<N>& x, const double* times, const long long n, const double dt,
                                  Obs& obs) {
  if (n == 0) return true;
  const double last = times[n - 1];
  long long it = 0;
  st.initialize(x, times[0], dt);
  obs(x, times[it++]);
  bool in_step = false, ok = true;
  bool done = (it == n);
  while (!done) {
    if (!in_step) {
      while (it != n && less_eq_with_sign(times[it], st.t, st.dt)) {
        st.calc_state(times[it], x);
        obs(x, times[it]);
        ++it;
      }
      if (less_eq_with_sign(st.t + st.dt, last, st.dt)) {
        st.begin_step(sys);  // (Boost takes this step even when every sample has been emitted)
      } else if (it != n) {
        st.initialize(st.current_state(), st.t, last - st.t);  // land exactly on the last time point
        st.begin_step(sys);
      } else {
        break;
      }
      in_step = true;
    }
    const int r = st.try_once(sys);
    if (r > 0) { in_step = false; done = (it == n); }
    else if (r < 0) { ok = false; done = true; }
  }
  return ok;
}

template <class St, class Sys, int N, class Obs>
ODEINTR_DEV bool flat_adaptive_dense(St& st, Sys& sys, vec<N>& x, const double t0, const double t1, const double dt,
                                     Obs& obs);

template <class St, class Sys, int N, class Obs>
ODEINTR_DEV bool flat_const_dense(St& st, Sys& sys, vec<N>& x, const double t0, const double t1, const double dt,
                                  Obs& obs) {
  double time = t0;
  st.initialize(x, time, dt);
  obs(x, time);
  time += dt;
  long long obs_step = 1;
  int mode = 0;  // 0: top of Boost's outer loop, 1: inside `while (st.t <= time) do_step`, 2: the single last step
  bool in_step = false, ok = true, done = false, stepped = false;
  int idle = 0;
  while (!done) {
    if (!in_step) {
      bool progressed = false;
      if (mode == 1) {
        if (less_eq_with_sign(st.t, time, dt)) { st.begin_step(sys); in_step = true; }
        else mode = 0;
      }
      if (mode == 0) {
        if (!less_eq_with_sign(time + dt, t1, dt)) break;
        while (less_eq_with_sign(time, st.t, dt)) {  // every sample the last real step already covers
          st.calc_state(time, x);
          obs(x, time);
          ++obs_step;
          time = t0 + static_cast<double>(obs_step) * dt;
          progressed = true;
        }
        if (less_with_sign(st.t + st.dt, t1, st.dt)) {
          if (less_eq_with_sign(st.t, time, dt)) { mode = 1; st.begin_step(sys); in_step = true; }
        } else if (less_with_sign(st.t, t1, st.dt)) {
          st.initialize(st.current_state(), st.t, t1 - st.t);
          mode = 2; st.begin_step(sys); in_step = true;
        }
      }
      if (!in_step) {
        // st.t is NaN (or the step collapsed): Boost's loop would spin here forever; stop this instance
        if (!progressed && ++idle > ODEINTR_NO_PROGRESS_LIMIT) {
          st.counters().status |= ODEINTR_STATUS_NONFINITE;
          return false;
        }
        continue;
      }
      idle = 0;
    }
    const int r = st.try_once(sys);
    if (r > 0) { in_step = false; stepped = true; if (mode == 2) mode = 0; }
    else if (r < 0) { ok = false; done = true; }
  }
  if (ok && less_eq_with_sign(time, t1, dt)) {
    if (!stepped) {
      // no step was taken (dt <= t1 - t0 < 2 dt), so there is no interpolant: integrate to the sample time with
      // real steps instead of Boost's read of unset stage vectors (see integrate_const_dense in integrate_loops.cuh)
      null_observer quiet;
      ok = flat_adaptive_dense(st, sys, x, t0, time, dt, quiet);
    } else {
      st.calc_state(time, x);
    }
    if (ok) obs(x, time);
  }
  return ok;
}

template <class St, class Sys, int N, class Obs>
ODEINTR_DEV bool flat_adaptive_dense(St& st, Sys& sys, vec<N>& x, const double t0, const double t1, const double dt,
                                     Obs& obs) {
  st.initialize(x, t0, dt);
  bool in_step = false, ok = true, done = false;
  bool inner = false;
  int idle = 0;
  while (!done) {
    if (!in_step) {
      if (!inner) {
        if (!less_with_sign(st.t, t1, st.dt)) break;
        inner = true;
      }
      if (less_eq_with_sign(st.t + st.dt, t1, st.dt)) {
        obs(st.current_state(), st.t);
        st.begin_step(sys);
        in_step = true;
        idle = 0;
      } else {
        st.initialize(st.current_state(), st.t, t1 - st.t);  // land exactly on t1
        inner = false;
        if (++idle > ODEINTR_NO_PROGRESS_LIMIT) { st.counters().status |= ODEINTR_STATUS_NONFINITE; return false; }
        continue;
      }
    }
    const int r = st.try_once(sys);
    if (r > 0) in_step = false;
    else if (r < 0) { ok = false; done = true; }
  }
  if (ok) {
    obs(st.current_state(), st.t);
    x = st.current_state();
  }
  return ok;
}

}  // namespace odeintr_b200
