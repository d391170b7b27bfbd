// odeintr_b200/csrc/kernels/rk78.cuh
// Runge-Kutta-Fehlberg 7(8): Boost's runge_kutta_fehlberg78 = explicit_error_generic_rk<13, 8, 8, 7>
// (methods "rk78" and "rk78_a", R/utils.R:53).  The 8th-order weights propagate the solution, the embedded
// error is dt * sum (b_i - b7_i) k_i.  Tableau-driven: the loops below are fully unrolled over constexpr
// coefficient arrays and zero entries are skipped at compile time, so what remains is Boost's left-to-right
// scale_sum over the non-zero terms with coefficient*dt formed first.
#pragma once
#include "odeintr_b200/adaptive_steppers.cuh"

namespace odeintr_b200 {

struct rkf78_tableau {
  static constexpr int S = 13;
  // a[i][j]: stage i+1 (0-based row i >= 1) uses k_{j+1}
  static constexpr double a[13][12] = {
      {0},
      {2.0 / 27.0},
      {1.0 / 36.0, 1.0 / 12.0},
      {1.0 / 24.0, 0.0, 1.0 / 8.0},
      {5.0 / 12.0, 0.0, -25.0 / 16.0, 25.0 / 16.0},
      {1.0 / 20.0, 0.0, 0.0, 1.0 / 4.0, 1.0 / 5.0},
      {-25.0 / 108.0, 0.0, 0.0, 125.0 / 108.0, -65.0 / 27.0, 125.0 / 54.0},
      {31.0 / 300.0, 0.0, 0.0, 0.0, 61.0 / 225.0, -2.0 / 9.0, 13.0 / 900.0},
      {2.0, 0.0, 0.0, -53.0 / 6.0, 704.0 / 45.0, -107.0 / 9.0, 67.0 / 90.0, 3.0},
      {-91.0 / 108.0, 0.0, 0.0, 23.0 / 108.0, -976.0 / 135.0, 311.0 / 54.0, -19.0 / 60.0, 17.0 / 6.0, -1.0 / 12.0},
      {2383.0 / 4100.0, 0.0, 0.0, -341.0 / 164.0, 4496.0 / 1025.0, -301.0 / 82.0, 2133.0 / 4100.0, 45.0 / 82.0,
       45.0 / 164.0, 18.0 / 41.0},
      {3.0 / 205.0, 0.0, 0.0, 0.0, 0.0, -6.0 / 41.0, -3.0 / 205.0, -3.0 / 41.0, 3.0 / 41.0, 6.0 / 41.0, 0.0},
      {-1777.0 / 4100.0, 0.0, 0.0, -341.0 / 164.0, 4496.0 / 1025.0, -289.0 / 82.0, 2193.0 / 4100.0, 51.0 / 82.0,
       33.0 / 164.0, 12.0 / 41.0, 0.0, 1.0}};
  static constexpr double c[13] = {0.0, 2.0 / 27.0, 1.0 / 9.0, 1.0 / 6.0, 5.0 / 12.0, 1.0 / 2.0, 5.0 / 6.0,
                                   1.0 / 6.0, 2.0 / 3.0, 1.0 / 3.0, 1.0, 0.0, 1.0};
  static constexpr double b[13] = {0.0, 0.0, 0.0, 0.0, 0.0, 34.0 / 105.0, 9.0 / 35.0, 9.0 / 35.0, 9.0 / 280.0,
                                   9.0 / 280.0, 0.0, 41.0 / 840.0, 41.0 / 840.0};
  static constexpr double db[13] = {0.0 - 41.0 / 840.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
                                    0.0 - 41.0 / 840.0, 41.0 / 840.0, 41.0 / 840.0};
};

// one step from (x, k1 = f(x, t)): xnew (8th order) and, if WITH_ERR, xerr
template <class Sys, int N, bool WITH_ERR>
ODEINTR_DEV void rkf78_step(Sys& sys, const vec<N>& x, const vec<N>& k1, const double t, const double dt, vec<N>& xnew,
                            vec<N>& xerr) {
  typedef rkf78_tableau T;
  vec<N> k[T::S];
  k[0] = k1;
#pragma unroll
  for (int s = 1; s < T::S; ++s) {
    vec<N> xt = x;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      if (T::a[s][j] != 0.0) {
        const double e = T::a[s][j] * dt;
#pragma unroll
        for (int i = 0; i < N; ++i) xt[i] = xt[i] + e * k[j][i];
      }
    }
    sys.sys(xt, k[s], t + T::c[s] * dt);
  }
  xnew = x;
#pragma unroll
  for (int j = 0; j < T::S; ++j) {
    if (T::b[j] != 0.0) {
      const double e = T::b[j] * dt;
#pragma unroll
      for (int i = 0; i < N; ++i) xnew[i] = xnew[i] + e * k[j][i];
    }
  }
  if (WITH_ERR) {
    bool first = true;
#pragma unroll
    for (int j = 0; j < T::S; ++j) {
      if (T::db[j] != 0.0) {
        const double e = T::db[j] * dt;
#pragma unroll
        for (int i = 0; i < N; ++i) xerr[i] = first ? e * k[j][i] : xerr[i] + e * k[j][i];
        first = false;
      }
    }
  }
}

template <class Sys, int N>
struct rk78_stepper {
  ODEINTR_DEV void reset() {}
  ODEINTR_DEV void do_step(Sys& sys, vec<N>& x, const double t, const double dt) {
    vec<N> k1, xn, xe;
    sys.sys(x, k1, t);
    rkf78_step<Sys, N, false>(sys, x, k1, t, dt, xn, xe);
    x = xn;
  }
};

// controlled_runge_kutta< runge_kutta_fehlberg78 >: error order 7, stepper order 8
template <class Sys, int N>
struct rk78_controlled {
  static constexpr bool is_dense = false;
  double eps_abs, eps_rel;
  step_counters cnt;
  ODEINTR_DEV void setup(const double atol, const double rtol) { eps_abs = atol; eps_rel = rtol; cnt.reset(); }
  ODEINTR_DEV void restart() {}
  ODEINTR_DEV step_counters& counters() { return cnt; }
  ODEINTR_DEV bool try_step(Sys& sys, vec<N>& x, double& t, double& dt) {
    vec<N> k1, xn, xe;
    sys.sys(x, k1, t);
    rkf78_step<Sys, N, true>(sys, x, k1, t, dt, xn, xe);
    const double err = reduce_error(sys, error_norm<N>(x, k1, xe, dt, eps_abs, eps_rel), 0);
    const double dt_used = dt;
    if (!adjust_step(err, dt, 7, 8, 2.56e-6)) {
      cnt.rejected++;
      return false;
    }
    t += dt_used;
    x = xn;
    cnt.accepted++;
    return true;
  }
};

}  // namespace odeintr_b200
