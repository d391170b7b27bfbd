// odeintr_b200/csrc/kernels/steppers_plain.cuh
// Fixed-step steppers, one trajectory per thread, state and stage derivatives in registers.
// They replace Boost's euler / runge_kutta4 / runge_kutta_cash_karp54 / runge_kutta_dopri5 used as
// plain steppers, selected by R/utils.R:48-53 of the reference.  The floating-point operation order
// is Boost's: coefficient*dt is formed first, sums run left to right, no
// FMA contraction (the TU is compiled with --fmad=false in exact mode).
//
// Boost's generic Runge-Kutta algorithm also adds the zero tableau entries (x + (0*dt)*k1 + ...).
// For finite k those terms add +/-0.0, which changes nothing except the sign of an exactly-zero
// state, so they are dropped here unless ODEINTR_KEEP_ZERO_TERMS is defined.
#pragma once
#include "odeintr_b200/prelude.cuh"

namespace odeintr_b200 {

template <class Sys, int N>
struct euler_stepper {
  ODEINTR_DEV void reset() {}
  ODEINTR_DEV void do_step(Sys& sys, vec<N>& x, const double t, const double dt) {
    vec<N> dxdt;
    sys.sys(x, dxdt, t);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = x[i] + dt * dxdt[i];
  }
};

// runge_kutta4<state_type> == explicit_generic_rk<4,4,...> with the classic tableau
template <class Sys, int N>
struct rk4_stepper {
  ODEINTR_DEV void reset() {}
  ODEINTR_DEV void do_step(Sys& sys, vec<N>& x, const double t, const double dt) {
    const double half = 1.0 / 2.0;
    const double b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
    vec<N> k1, k2, k3, k4, xt;
    sys.sys(x, k1, t);
    {
      const double a0 = half * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = x[i] + a0 * k1[i];
    }
    sys.sys(xt, k2, t + half * dt);
    {
      const double a1 = half * dt;
#ifdef ODEINTR_KEEP_ZERO_TERMS
      const double a0 = 0.0 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = x[i] + a0 * k1[i] + a1 * k2[i];
#else
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = x[i] + a1 * k2[i];
#endif
    }
    sys.sys(xt, k3, t + half * dt);
    {
      const double a2 = 1.0 * dt;
#ifdef ODEINTR_KEEP_ZERO_TERMS
      const double a0 = 0.0 * dt, a1 = 0.0 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = x[i] + a0 * k1[i] + a1 * k2[i] + a2 * k3[i];
#else
#pragma unroll
      for (int i = 0; i < N; ++i) xt[i] = x[i] + a2 * k3[i];
#endif
    }
    sys.sys(xt, k4, t + 1.0 * dt);
    {
      const double c0 = b1 * dt, c1 = b2 * dt, c2 = b2 * dt, c3 = b1 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] = x[i] + c0 * k1[i] + c1 * k2[i] + c2 * k3[i] + c3 * k4[i];
    }
  }
};

// runge_kutta_cash_karp54 used as a plain stepper (method "rk54"): 5th-order solution only.
template <class Sys, int N>
struct rk54_stepper {
  ODEINTR_DEV void reset() {}
  ODEINTR_DEV void do_step(Sys& sys, vec<N>& x, const double t, const double dt) {
    const double a21 = 1.0 / 5.0;
    const double a31 = 3.0 / 40.0, a32 = 9.0 / 40.0;
    const double a41 = 3.0 / 10.0, a42 = -9.0 / 10.0, a43 = 6.0 / 5.0;
    const double a51 = -11.0 / 54.0, a52 = 5.0 / 2.0, a53 = -70.0 / 27.0, a54 = 35.0 / 27.0;
    const double a61 = 1631.0 / 55296.0, a62 = 175.0 / 512.0, a63 = 575.0 / 13824.0,
                 a64 = 44275.0 / 110592.0, a65 = 253.0 / 4096.0;
    const double b1 = 37.0 / 378.0, b3 = 250.0 / 621.0, b4 = 125.0 / 594.0, b6 = 512.0 / 1771.0;
    const double c2 = 1.0 / 5.0, c3 = 3.0 / 10.0, c4 = 3.0 / 5.0, c5 = 1.0, c6 = 7.0 / 8.0;
    vec<N> k1, k2, k3, k4, k5, k6, xt;
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
    { const double e1 = b1 * dt, e3 = b3 * dt, e4 = b4 * dt, e6 = b6 * dt;
#ifdef ODEINTR_KEEP_ZERO_TERMS
      const double e2 = 0.0 * dt, e5 = 0.0 * dt;
#pragma unroll
      for (int i = 0; i < N; ++i)
        x[i] = x[i] + e1 * k1[i] + e2 * k2[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i];
#else
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] = x[i] + e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e6 * k6[i];
#endif
    }
  }
};

}  // namespace odeintr_b200
