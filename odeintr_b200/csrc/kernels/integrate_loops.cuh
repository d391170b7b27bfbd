// odeintr_b200/csrc/kernels/integrate_loops.cuh
// Boost's integrate_const / integrate_times / integrate_adaptive for controlled and dense-output steppers
// (boost/numeric/odeint/integrate/detail/*.hpp; control flow restated in SURVEY.md Appendix A.3-A.5), written
// once for both users: every GPU thread runs these loops for its own trajectory (State = register vector,
// integrate_device.cuh), and the host driver of large systems runs them over device-resident vectors
// (State = lattice vector handle, csrc/lattice_adaptive.h).  All loops return false when the stepper
// reported a failure (Boost's 500-failed-steps overflow, the step budget, no progress).
#pragma once
#include "odeintr_b200/time_compare.cuh"

namespace odeintr_b200 {

#ifndef ODEINTR_MAX_FAILED_STEPS
#define ODEINTR_MAX_FAILED_STEPS 500  // Boost failed_step_checker
#endif
// Boost's do { try_step; fail_checker(); } while (fail) loops call the checker after EVERY attempt, so the 501st
// attempt of a step throws whatever its outcome: a step may use at most 500 attempts (`++fails >= 500` after a
// rejection).  integrate_times' controlled loop calls the checker on rejections only: 501 of them (`++fails > 500`).

struct null_observer {
  template <class State> ODEINTR_HD void operator()(const State&, double) const {}
};

// ---- integrate_adaptive --------------------------------------------------------------------------------
#pragma nv_exec_check_disable
template <class St, class Sys, class State, class Obs>
ODEINTR_HD bool integrate_adaptive_controlled(St& st, Sys& sys, State& x, double& t, const double t1, double& dt,
                                               Obs& obs) {
  int fails = 0;
  while (less_with_sign(t, t1, dt)) {
    obs(x, t);
    if (less_with_sign(t1, t + dt, dt)) dt = t1 - t;
    if (st.counters().exhausted()) return false;
    while (!st.try_step(sys, x, t, dt)) {
      if (++fails >= ODEINTR_MAX_FAILED_STEPS) { st.counters().status |= ODEINTR_STATUS_STEP_OVERFLOW; return false; }
    }
    fails = 0;
  }
  obs(x, t);
  return true;
}

#pragma nv_exec_check_disable
template <class St, class Sys, class State, class Obs>
ODEINTR_HD bool integrate_adaptive_dense(St& st, Sys& sys, State& x, const double t0, const double t1,
                                          const double dt, Obs& obs) {
  st.initialize(x, t0, dt);
  while (less_with_sign(st.t, t1, st.dt)) {
    while (less_eq_with_sign(st.t + st.dt, t1, st.dt)) {
      obs(st.current_state(), st.t);
      if (!st.do_step(sys)) return false;
    }
    st.initialize(st.current_state(), st.t, t1 - st.t);  // land exactly on t1
  }
  obs(st.current_state(), st.t);
  x = st.current_state();
  return true;
}

// ---- integrate_const -----------------------------------------------------------------------------------
#pragma nv_exec_check_disable
template <class St, class Sys, class State, class Obs>
ODEINTR_HD bool integrate_const_controlled(St& st, Sys& sys, State& x, const double t0, const double t1, double dt,
                                            Obs& obs) {
  double time = t0;
  const double time_step = dt;
  long long step = 0;
  null_observer nobs;
  while (less_eq_with_sign(time + time_step, t1, dt)) {
    obs(x, time);
    double tcur = time;
    st.restart();  // Boost hands every interval a fresh copy of the stepper (FSAL derivative re-evaluated)
    if (!integrate_adaptive_controlled(st, sys, x, tcur, time + time_step, dt, nobs)) return false;
    ++step;
    time = t0 + static_cast<double>(step) * time_step;
  }
  obs(x, time);
  return true;
}

#pragma nv_exec_check_disable
template <class St, class Sys, class State, class Obs>
ODEINTR_HD bool integrate_const_dense(St& st, Sys& sys, State& x, const double t0, const double t1, const double dt,
                                       Obs& obs) {
  double time = t0;
  st.initialize(x, time, dt);
  obs(x, time);
  time += dt;
  long long obs_step = 1;
  bool stepped = false;
  while (less_eq_with_sign(time + dt, t1, dt)) {
    bool progressed = false;
    while (less_eq_with_sign(time, st.t, dt)) {  // every sample the last real step already covers
      st.calc_state(time, x);
      obs(x, time);
      ++obs_step;
      time = t0 + static_cast<double>(obs_step) * dt;
      progressed = true;
    }
    if (less_with_sign(st.t + st.dt, t1, st.dt)) {
      while (less_eq_with_sign(st.t, time, dt)) {
        if (!st.do_step(sys)) return false;
        progressed = true; stepped = true;
      }
    } else if (less_with_sign(st.t, t1, st.dt)) {
      st.initialize(st.current_state(), st.t, t1 - st.t);
      if (!st.do_step(sys)) return false;
      progressed = true; stepped = true;
    }
    if (!progressed) {
      // st.t is NaN (or the step collapsed): Boost's loop would spin here forever; stop this instance
      st.counters().status |= ODEINTR_STATUS_NONFINITE;
      return false;
    }
  }
  if (less_eq_with_sign(time, t1, dt)) {
    if (!stepped) {
      // dt <= t1 - t0 < 2 dt: the loop above never took a step, so there is no interpolant; Boost calls
      // calc_state anyway and reads unset stage vectors (undefined behaviour - this is what name_at() runs into
      // when times[0] - start equals step_size, e.g. the documented bistable_at(inic, 10^(0:3)) example).
      // Conscious fix: integrate to the sample time with real steps instead.
      null_observer quiet;
      if (!integrate_adaptive_dense(st, sys, x, t0, time, dt, quiet)) return false;
    } else {
      st.calc_state(time, x);
    }
    obs(x, time);
  }
  return true;
}

// ---- integrate_times -----------------------------------------------------------------------------------
#pragma nv_exec_check_disable
template <class St, class Sys, class State, class Obs>
ODEINTR_HD bool integrate_times_controlled(St& st, Sys& sys, State& x, const double* times, const long long n,
                                            double dt, Obs& obs) {
  if (n == 0) return true;
  int fails = 0;
  long long it = 0;
  while (true) {
    double current_time = times[it++];
    obs(x, current_time);
    if (it == n) break;
    const double next = times[it];
    while (less_with_sign(current_time, next, dt)) {
      double current_dt = dt > 0 ? fmin(dt, next - current_time) : fmax(dt, next - current_time);  // min_abs
      if (st.counters().exhausted()) return false;
      if (st.try_step(sys, x, current_time, current_dt)) {
        fails = 0;
        dt = dt > 0 ? fmax(dt, current_dt) : fmin(dt, current_dt);  // max_abs: keep the larger step
      } else {
        if (++fails > ODEINTR_MAX_FAILED_STEPS) { st.counters().status |= ODEINTR_STATUS_STEP_OVERFLOW; return false; }
        dt = current_dt;
      }
    }
  }
  return true;
}

#pragma nv_exec_check_disable
template <class St, class Sys, class State, class Obs>
ODEINTR_HD bool integrate_times_dense(St& st, Sys& sys, State& x, const double* times, const long long n,
                                       const double dt, Obs& obs) {
  if (n == 0) return true;
  const double last = times[n - 1];
  long long it = 0;
  st.initialize(x, times[0], dt);
  obs(x, times[it++]);
  while (it != n) {
    while (it != n && less_eq_with_sign(times[it], st.t, st.dt)) {
      st.calc_state(times[it], x);
      obs(x, times[it]);
      ++it;
    }
    if (less_eq_with_sign(st.t + st.dt, last, st.dt)) {
      if (!st.do_step(sys)) return false;
    } else if (it != n) {
      st.initialize(st.current_state(), st.t, last - st.t);  // land exactly on the last time point
      if (!st.do_step(sys)) return false;
    }
  }
  return true;
}

}  // namespace odeintr_b200
