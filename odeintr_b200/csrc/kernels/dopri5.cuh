// odeintr_b200/csrc/kernels/dopri5.cuh
// Dormand-Prince 5(4): step, embedded error, step-size controller and dense output, one trajectory
// per thread with every stage derivative in registers.  Replaces Boost's runge_kutta_dopri5,
// controlled_runge_kutta<> (default_error_checker + default_step_adjuster) and
// dense_output_runge_kutta<> selected by the reference at R/utils.R:27-30,52 (methods rk5, rk5_a,
// rk5_i = the default).  Operation order follows SURVEY.md Appendix A.2.
#pragma once
#include "odeintr_b200/prelude.cuh"

namespace odeintr_b200 {

// The tableau lives in constant memory: an FP64 instruction takes a constant-bank operand directly, whereas a literal
// costs two UMOV per use (ncu on config 3: 15 % of the issued instructions were UMOVs re-forming these constants).
// Same doubles as the literals they replace (the initialisers are the same constant expressions).
#ifndef ODEINTR_DP5_LITERALS
__constant__ double dp5_tab[] = {
    1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0,                                                     // a2 .. a5
    1.0 / 5.0,                                                                                         // b21
    3.0 / 40.0, 9.0 / 40.0,                                                                            // b31, b32
    44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0,                                                             // b41 .. b43
    19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0,                             // b51 .. b54
    9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0,                 // b61 .. b65
    35.0 / 384.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0,                        // c1, c3 .. c6
    35.0 / 384.0 - 5179.0 / 57600.0, 500.0 / 1113.0 - 7571.0 / 16695.0, 125.0 / 192.0 - 393.0 / 640.0,  // dc1, dc3, dc4
    -2187.0 / 6784.0 - (-92097.0 / 339200.0), 11.0 / 84.0 - 187.0 / 2100.0, -1.0 / 40.0};              // dc5, dc6, dc7
#define ODEINTR_DP5(i, literal) dp5_tab[i]
#else
#define ODEINTR_DP5(i, literal) (literal)
#endif

template <class Sys, int N>
struct dopri5_core {
  // stage derivatives of the last computed step (k2 is not needed after the step: b2 = 0)
  vec<N> k3, k4, k5, k6;

  // out = in + dt * sum b_i k_i ; dxdt_out = f(out, t + dt)   (FSAL)
  ODEINTR_DEV void do_step_impl(Sys& sys, const vec<N>& in, const vec<N>& dxdt_in, const double t, vec<N>& out,
                                vec<N>& dxdt_out, const double dt) {
    const double a2 = ODEINTR_DP5(0, 1.0 / 5.0), a3 = ODEINTR_DP5(1, 3.0 / 10.0), a4 = ODEINTR_DP5(2, 4.0 / 5.0),
                 a5 = ODEINTR_DP5(3, 8.0 / 9.0);
    const double b21 = ODEINTR_DP5(4, 1.0 / 5.0);
    const double b31 = ODEINTR_DP5(5, 3.0 / 40.0), b32 = ODEINTR_DP5(6, 9.0 / 40.0);
    const double b41 = ODEINTR_DP5(7, 44.0 / 45.0), b42 = ODEINTR_DP5(8, -56.0 / 15.0), b43 = ODEINTR_DP5(9, 32.0 / 9.0);
    const double b51 = ODEINTR_DP5(10, 19372.0 / 6561.0), b52 = ODEINTR_DP5(11, -25360.0 / 2187.0),
                 b53 = ODEINTR_DP5(12, 64448.0 / 6561.0), b54 = ODEINTR_DP5(13, -212.0 / 729.0);
    const double b61 = ODEINTR_DP5(14, 9017.0 / 3168.0), b62 = ODEINTR_DP5(15, -355.0 / 33.0),
                 b63 = ODEINTR_DP5(16, 46732.0 / 5247.0), b64 = ODEINTR_DP5(17, 49.0 / 176.0),
                 b65 = ODEINTR_DP5(18, -5103.0 / 18656.0);
    const double c1 = ODEINTR_DP5(19, 35.0 / 384.0), c3 = ODEINTR_DP5(20, 500.0 / 1113.0), c4 = ODEINTR_DP5(21, 125.0 / 192.0),
                 c5 = ODEINTR_DP5(22, -2187.0 / 6784.0), c6 = ODEINTR_DP5(23, 11.0 / 84.0);
    vec<N> xt, k2;
    { const double e1 = dt * b21;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = in[i] + e1 * dxdt_in[i]; }
    sys.sys(xt, k2, t + dt * a2);
    { const double e1 = dt * b31, e2 = dt * b32;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = in[i] + e1 * dxdt_in[i] + e2 * k2[i]; }
    sys.sys(xt, k3, t + dt * a3);
    { const double e1 = dt * b41, e2 = dt * b42, e3 = dt * b43;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = in[i] + e1 * dxdt_in[i] + e2 * k2[i] + e3 * k3[i]; }
    sys.sys(xt, k4, t + dt * a4);
    { const double e1 = dt * b51, e2 = dt * b52, e3 = dt * b53, e4 = dt * b54;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = in[i] + e1 * dxdt_in[i] + e2 * k2[i] + e3 * k3[i] + e4 * k4[i]; }
    sys.sys(xt, k5, t + dt * a5);
    { const double e1 = dt * b61, e2 = dt * b62, e3 = dt * b63, e4 = dt * b64, e5 = dt * b65;
#pragma unroll
      for (int i = 0; i < N; ++i)
        xt[i] = in[i] + e1 * dxdt_in[i] + e2 * k2[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i]; }
    sys.sys(xt, k6, t + dt);
    { const double e1 = dt * c1, e3 = dt * c3, e4 = dt * c4, e5 = dt * c5, e6 = dt * c6;
#pragma unroll
      for (int i = 0; i < N; ++i)
        out[i] = in[i] + e1 * dxdt_in[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i]; }
    sys.sys(out, dxdt_out, t + dt);
  }

  // embedded error estimate x_err = dt * sum dc_i k_i (k7 = dxdt_out)
  ODEINTR_DEV void calc_error(const vec<N>& dxdt_in, const vec<N>& dxdt_out, const double dt, vec<N>& xerr) const {
    const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0,
                 c6 = 11.0 / 84.0;
    const double dc1 = ODEINTR_DP5(24, c1 - 5179.0 / 57600.0), dc3 = ODEINTR_DP5(25, c3 - 7571.0 / 16695.0),
                 dc4 = ODEINTR_DP5(26, c4 - 393.0 / 640.0), dc5 = ODEINTR_DP5(27, c5 - (-92097.0 / 339200.0)),
                 dc6 = ODEINTR_DP5(28, c6 - 187.0 / 2100.0), dc7 = ODEINTR_DP5(29, -1.0 / 40.0);
    (void)c1; (void)c3; (void)c4; (void)c5; (void)c6;
    const double e1 = dt * dc1, e3 = dt * dc3, e4 = dt * dc4, e5 = dt * dc5, e6 = dt * dc6, e7 = dt * dc7;
#pragma unroll
    for (int i = 0; i < N; ++i)
      xerr[i] = e1 * dxdt_in[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * dxdt_out[i];
  }

  // 4th-order dense output between (t_old, x_old, deriv_old) and (t_new, ., deriv_new)
  ODEINTR_DEV void calc_state(const double t, vec<N>& x, const vec<N>& x_old, const vec<N>& deriv_old,
                              const double t_old, const vec<N>& deriv_new, const double t_new) const {
    const double b1 = 35.0 / 384.0, b3 = 500.0 / 1113.0, b4 = 125.0 / 192.0, b5 = -2187.0 / 6784.0,
                 b6 = 11.0 / 84.0;
    const double dt = (t_new - t_old);
    const double theta = (t - t_old) / dt;
    // Boost divides by these six constants; the quotients below are the same bits (exact_div.cuh) at a third of the cost
    constexpr divisor c1(11282082432.0, 1.0 / 11282082432.0), c3(32700410799.0, 1.0 / 32700410799.0),
        c4(1880347072.0, 1.0 / 1880347072.0), c5(199316789632.0, 1.0 / 199316789632.0),
        c6(822651844.0, 1.0 / 822651844.0), c7(29380423.0, 1.0 / 29380423.0);
    const double X1 = c1.under(5.0 * (2558722523.0 - 31403016.0 * theta));
    const double X3 = c3.under(100.0 * (882725551.0 - 15701508.0 * theta));
    const double X4 = c4.under(25.0 * (443332067.0 - 31403016.0 * theta));
    const double X5 = c5.under(32805.0 * (23143187.0 - 3489224.0 * theta));
    const double X6 = c6.under(55.0 * (29972135.0 - 7076736.0 * theta));
    const double X7 = c7.under(10.0 * (7414447.0 - 829305.0 * theta));
    const double theta_m_1 = theta - 1.0;
    const double theta_sq = theta * theta;
    const double A = theta_sq * (3.0 - 2.0 * theta);
    const double B = theta_sq * theta_m_1;
    const double C = theta_sq * theta_m_1 * theta_m_1;
    const double D = theta * theta_m_1 * theta_m_1;
    const double b1_theta = A * b1 - C * X1 + D;
    const double b3_theta = A * b3 + C * X3;
    const double b4_theta = A * b4 - C * X4;
    const double b5_theta = A * b5 + C * X5;
    const double b6_theta = A * b6 - C * X6;
    const double b7_theta = B + C * X7;
    const double e1 = dt * b1_theta, e3 = dt * b3_theta, e4 = dt * b4_theta, e5 = dt * b5_theta,
                 e6 = dt * b6_theta, e7 = dt * b7_theta;
#pragma unroll
    for (int i = 0; i < N; ++i)
      x[i] = x_old[i] + e1 * deriv_old[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * deriv_new[i];
  }
};

// default_error_checker::error: max_i |xerr_i| / (eps_abs + eps_rel * (|x_i| + |dt| |dxdt_i|)), a_x = a_dxdt = 1
template <int N>
ODEINTR_DEV double error_norm(const vec<N>& x_old, const vec<N>& dxdt_old, const vec<N>& xerr, const double dt,
                              const double eps_abs, const double eps_rel) {
  const double adt = 1.0 * fabs(dt);
  double m = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double e = fabs(xerr[i]) / (eps_abs + eps_rel * (1.0 * fabs(x_old[i]) + adt * fabs(dxdt_old[i])));
    m = fmax(m, e);
  }
  return m;
}

// default_step_adjuster (max_dt = 0): decrease_step on rejection, increase_step on acceptance
//   reject:  dt *= max(0.9 * pow(err, -1/(error_order-1)), 1/5)
//   accept:  if err < 0.5:  dt *= 0.9 * pow(max(5^-stepper_order, err), -1/stepper_order)
// folded into ONE pow call per attempt with per-lane arguments: the values are exactly Boost's, but the lanes of a
// warp that reject, grow or keep their step no longer execute two serialized pow expansions (~100 instructions
// each).  5^-stepper_order is passed as the literal glibc's pow returns (5^-5 = 0.00032, 5^-8 = 2.56e-6).
// Returns true when the step is accepted; the caller advances t with the OLD dt before calling.
ODEINTR_DEV bool adjust_step(const double err, double& dt, const int error_order, const int stepper_order,
                             const double pow5_minus_order) {
  const bool rejected = err > 1.0;
  const double base = rejected ? err : fmax(pow5_minus_order, err);
  const double expo = rejected ? -1.0 / (error_order - 1) : -1.0 / stepper_order;
  // (no lane of the warp needs it when every error lies in [0.5, 1]: common in smooth, sorted sweeps)
  const bool needed = rejected || err < 0.5;
  double p = 1.0;
  if (__any_sync(__activemask(), needed)) p = pow(base, expo);
  if (rejected) dt *= fmax(9.0 / 10.0 * p, 1.0 / 5.0);
  else if (err < 0.5) dt *= 9.0 / 10.0 * p;
  return !rejected;
}

// runge_kutta_dopri5 as a plain stepper (method "rk5"): explicit_error_stepper_fsal_base::do_step keeps
// the derivative of the previous step and evaluates it only on the first call.
template <class Sys, int N>
struct rk5_stepper {
  dopri5_core<Sys, N> core;
  vec<N> dxdt;
  bool first;
  ODEINTR_DEV void reset() { first = true; }
  ODEINTR_DEV void do_step(Sys& sys, vec<N>& x, const double t, const double dt) {
    if (first) { sys.sys(x, dxdt, t); first = false; }
    vec<N> xn, dn;
    core.do_step_impl(sys, x, dxdt, t, xn, dn, dt);
    x = xn; dxdt = dn;
  }
};

}  // namespace odeintr_b200
