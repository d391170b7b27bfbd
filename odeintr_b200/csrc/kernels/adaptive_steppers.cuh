// odeintr_b200/csrc/kernels/adaptive_steppers.cuh
// Per-trajectory adaptive steppers with the interfaces the device integrate loops use:
//   controlled:  try_step(sys, x, t, dt) -> bool success      (Boost controlled_runge_kutta<>)
//   dense:       initialize / do_step / calc_state            (Boost dense_output_runge_kutta<>)
// State of one trajectory = a handful of register-resident vec<N>.  Operation order follows
// SURVEY.md Appendix A.2; every accepted / rejected try is counted per instance (north_star:
// accepted-step counts are reported next to the results).
#pragma once
#include "odeintr_b200/dopri5.cuh"

namespace odeintr_b200 {

#ifndef ODEINTR_MAX_FAILED_STEPS
#define ODEINTR_MAX_FAILED_STEPS 500  // Boost failed_step_checker
#endif

struct step_counters {
  long long accepted, rejected, budget;
  int status;
  ODEINTR_HD void reset() { accepted = 0; rejected = 0; status = ODEINTR_STATUS_OK; budget = 0x7fffffffffffffffLL; }
  // true (and the status is set) once the instance has used up its step-attempt budget
  ODEINTR_HD bool exhausted() {
    if (accepted + rejected < budget) return false;
    status |= ODEINTR_STATUS_STEP_BUDGET;
    return true;
  }
};

// A system whose state vector is spread over the threads of a trajectory (lattice_ensemble_adaptive.cuh) completes the
// error norm across them; an ordinary system's norm is complete as it is.
template <class Sys> ODEINTR_DEV auto reduce_error(Sys& s, const double e, int) -> decltype(s.block_max(e)) { return s.block_max(e); }
template <class Sys> ODEINTR_DEV double reduce_error(Sys&, const double e, long) { return e; }

// controlled_runge_kutta< runge_kutta_dopri5 > (FSAL): method "rk5_a"
template <class Sys, int N>
struct dopri5_controlled {
  static constexpr bool is_dense = false;
  dopri5_core<Sys, N> core;
  vec<N> dxdt;
  bool first;
  double eps_abs, eps_rel;
  step_counters cnt;
  ODEINTR_DEV void setup(const double atol, const double rtol) { eps_abs = atol; eps_rel = rtol; first = true; cnt.reset(); }
  ODEINTR_DEV void restart() { first = true; }
  ODEINTR_DEV step_counters& counters() { return cnt; }

  // try_step(sys, in, dxdt_in, t, out, dxdt_out, dt): t and dt are updated as Boost does
  // (measured, profiles/r02_bench_outline.log: one out-of-line copy of the attempt instead of one per integrate loop
  //  shrinks the kernel but passes the vectors through local memory: 4.23e10 -> 2.84e10 accepted steps/s; kept inline)
  ODEINTR_DEV bool try_step_full(Sys& sys, const vec<N>& in, const vec<N>& dxdt_in, double& t, vec<N>& out,
                                 vec<N>& dxdt_out, double& dt) {
    vec<N> xerr;
    core.do_step_impl(sys, in, dxdt_in, t, out, dxdt_out, dt);
    core.calc_error(dxdt_in, dxdt_out, dt, xerr);
    const double err = reduce_error(sys, error_norm<N>(in, dxdt_in, xerr, dt, eps_abs, eps_rel), 0);
    const double dt_used = dt;
    if (!adjust_step(err, dt, 4, 5, 0.00032)) {
      cnt.rejected++;
      return false;
    }
    t += dt_used;
    cnt.accepted++;
    return true;
  }
  ODEINTR_DEV bool try_step(Sys& sys, vec<N>& x, double& t, double& dt) {
    if (first) { sys.sys(x, dxdt, t); first = false; }
    vec<N> xn, dn;
    const bool ok = try_step_full(sys, x, dxdt, t, xn, dn, dt);
    if (ok) { x = xn; dxdt = dn; }
    return ok;
  }
};

// controlled_runge_kutta< runge_kutta_cash_karp54 >: method "rk54_a"
template <class Sys, int N>
struct rk54_controlled {
  static constexpr bool is_dense = false;
  double eps_abs, eps_rel;
  step_counters cnt;
  ODEINTR_DEV void setup(const double atol, const double rtol) { eps_abs = atol; eps_rel = rtol; cnt.reset(); }
  ODEINTR_DEV void restart() {}
  ODEINTR_DEV step_counters& counters() { return cnt; }
  ODEINTR_DEV bool try_step(Sys& sys, vec<N>& x, double& t, double& dt) {
    const double a21 = 1.0 / 5.0;
    const double a31 = 3.0 / 40.0, a32 = 9.0 / 40.0;
    const double a41 = 3.0 / 10.0, a42 = -9.0 / 10.0, a43 = 6.0 / 5.0;
    const double a51 = -11.0 / 54.0, a52 = 5.0 / 2.0, a53 = -70.0 / 27.0, a54 = 35.0 / 27.0;
    const double a61 = 1631.0 / 55296.0, a62 = 175.0 / 512.0, a63 = 575.0 / 13824.0,
                 a64 = 44275.0 / 110592.0, a65 = 253.0 / 4096.0;
    const double b1 = 37.0 / 378.0, b3 = 250.0 / 621.0, b4 = 125.0 / 594.0, b6 = 512.0 / 1771.0;
    const double db1 = b1 - 2825.0 / 27648.0, db3 = b3 - 18575.0 / 48384.0, db4 = b4 - 13525.0 / 55296.0,
                 db5 = -277.0 / 14336.0, db6 = b6 - 1.0 / 4.0;
    const double c2 = 1.0 / 5.0, c3 = 3.0 / 10.0, c4 = 3.0 / 5.0, c5 = 1.0, c6 = 7.0 / 8.0;
    vec<N> k1, k2, k3, k4, k5, k6, xt, xerr, xnew;
    sys.sys(x, k1, t);
    { const double e1 = a21 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = x[i] + e1 * k1[i]; }
    sys.sys(xt, k2, t + c2 * dt);
    { const double e1 = a31 * dt, e2 = a32 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = x[i] + e1 * k1[i] + e2 * k2[i]; }
    sys.sys(xt, k3, t + c3 * dt);
    { const double e1 = a41 * dt, e2 = a42 * dt, e3 = a43 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = x[i] + e1 * k1[i] + e2 * k2[i] + e3 * k3[i]; }
    sys.sys(xt, k4, t + c4 * dt);
    { const double e1 = a51 * dt, e2 = a52 * dt, e3 = a53 * dt, e4 = a54 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = x[i] + e1 * k1[i] + e2 * k2[i] + e3 * k3[i] + e4 * k4[i]; }
    sys.sys(xt, k5, t + c5 * dt);
    { const double e1 = a61 * dt, e2 = a62 * dt, e3 = a63 * dt, e4 = a64 * dt, e5 = a65 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i)
        xt[i] = x[i] + e1 * k1[i] + e2 * k2[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i]; }
    sys.sys(xt, k6, t + c6 * dt);
    { // Boost's generic algorithm forms the error from the full weight row (zero entries add +/-0)
      const double f1 = db1 * dt, f3 = db3 * dt, f4 = db4 * dt, f5 = db5 * dt, f6 = db6 * dt;
      const double e1 = b1 * dt, e3 = b3 * dt, e4 = b4 * dt, e6 = b6 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        xerr[i] = f1 * k1[i] + f3 * k3[i] + f4 * k4[i] + f5 * k5[i] + f6 * k6[i];
        xnew[i] = x[i] + e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e6 * k6[i];
      }
    }
    const double err = reduce_error(sys, error_norm<N>(x, k1, xerr, dt, eps_abs, eps_rel), 0);
    const double dt_used = dt;
    if (!adjust_step(err, dt, 4, 5, 0.00032)) {
      cnt.rejected++;
      return false;
    }
    t += dt_used;
    x = xnew;
    cnt.accepted++;
    return true;
  }
};

// dense_output_runge_kutta< controlled_runge_kutta< runge_kutta_dopri5 > >: method "rk5_i" (the default)
template <class Sys, int N>
struct dopri5_dense {
  static constexpr bool is_dense = true;
  dopri5_controlled<Sys, N> ctl;
  vec<N> x_cur, d_cur, x_old, d_old;
  double t, t_old, dt;
  bool deriv_ready;
  ODEINTR_DEV void setup(const double atol, const double rtol) { ctl.setup(atol, rtol); }
  ODEINTR_DEV step_counters& counters() { return ctl.cnt; }
  ODEINTR_DEV void initialize(const vec<N>& x0, const double t0, const double dt0) {
    x_cur = x0; t = t0; dt = dt0; deriv_ready = false;
  }
  ODEINTR_DEV const vec<N>& current_state() const { return x_cur; }
  // one accepted step; false = 500 consecutive rejections (Boost throws step_adjustment_error)
  ODEINTR_DEV bool do_step(Sys& sys) {
    if (!deriv_ready) { sys.sys(x_cur, d_cur, t); deriv_ready = true; }
    t_old = t;
    vec<N> xn, dn;
    int fails = 0;
    if (ctl.cnt.exhausted()) return false;
    while (!ctl.try_step_full(sys, x_cur, d_cur, t, xn, dn, dt)) {
      if (++fails >= ODEINTR_MAX_FAILED_STEPS) { ctl.cnt.status |= ODEINTR_STATUS_STEP_OVERFLOW; return false; }
    }
    x_old = x_cur; d_old = d_cur;
    x_cur = xn; d_cur = dn;
    return true;
  }
  ODEINTR_DEV void calc_state(const double tq, vec<N>& x) const {
    ctl.core.calc_state(tq, x, x_old, d_old, t_old, d_cur, t);
  }

  // do_step split into "start a step" + "one attempt", for the warp-synchronous loops of integrate_device.cuh:
  // a lane whose attempt was rejected retries in the warp's next iteration together with the lanes that start
  // their next step, instead of making the whole warp wait inside a retry loop.
  int fails;
  ODEINTR_DEV void begin_step(Sys& sys) {
    if (!deriv_ready) { sys.sys(x_cur, d_cur, t); deriv_ready = true; }
    t_old = t;
    fails = 0;
  }
  // 1 = step accepted, 0 = rejected (dt was reduced; call again), -1 = failed (status says why)
  ODEINTR_DEV int try_once(Sys& sys) {
    if (ctl.cnt.exhausted()) return -1;
    vec<N> xn, dn;
    if (ctl.try_step_full(sys, x_cur, d_cur, t, xn, dn, dt)) {
      x_old = x_cur; d_old = d_cur;
      x_cur = xn; d_cur = dn;
      return 1;
    }
    if (++fails >= ODEINTR_MAX_FAILED_STEPS) { ctl.cnt.status |= ODEINTR_STATUS_STEP_OVERFLOW; return -1; }
    return 0;
  }
};

}  // namespace odeintr_b200
