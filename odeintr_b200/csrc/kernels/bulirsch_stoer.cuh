// odeintr_b200/csrc/kernels/bulirsch_stoer.cuh
// Bulirsch-Stoer controlled stepper, one trajectory per thread: Boost's bulirsch_stoer<state_type> over
// modified_midpoint (method "bs", R/utils.R:57).  The reference default-constructs it (`stepper_type()`), so the
// tolerances are 1e-6 / 1e-6 whatever atol / rtol were passed to compile_sys (R/utils.R:22-32).
// Extrapolation tableau rows (up to k_max = 8), the midpoint work vectors and the order / step-size controller
// state are per thread; the stage loop over k is unrolled so every tableau row is addressed statically.
#pragma once
#include "odeintr_b200/adaptive_steppers.cuh"

namespace odeintr_b200 {

template <class Sys, int N>
struct bs_controlled {
  static constexpr bool is_dense = false;
  static constexpr int K_MAX = 8;
  double eps_abs, eps_rel;
  step_counters cnt;
  bool last_step_rejected, first;
  double dt_last;
  int k_opt;
  vec<N> table[K_MAX];

  ODEINTR_DEV void setup(const double, const double) {
    eps_abs = 1e-6; eps_rel = 1e-6;  // the reference's stepper_type() ignores atol / rtol for bs
    cnt.reset();
    dt_last = 0.0;
    reset();
  }
  ODEINTR_DEV void reset() { first = true; last_step_rejected = false; k_opt = 4; }
  // integrate_const hands every observation interval a fresh copy of the stepper: order and step history restart
  ODEINTR_DEV void restart() { reset(); dt_last = 0.0; }
  ODEINTR_DEV step_counters& counters() { return cnt; }

  static ODEINTR_DEV double interval(const int i) { return static_cast<double>(2 * (i + 1)); }
  static ODEINTR_DEV double cost(const int i) { return static_cast<double>((i + 1) * (i + 2)); }  // sum of 2(j+1), j<=i

  // modified_midpoint::do_step_impl with `steps` substeps
  ODEINTR_DEV void midpoint(Sys& sys, const vec<N>& in, const vec<N>& dxdt, const double t, vec<N>& out, const double dt,
                            const int steps) {
    const double h = dt / static_cast<double>(steps);
    const double h2 = 2.0 * h;
    double th = t + h;
    vec<N> x0, x1, d;
#pragma unroll
    for (int i = 0; i < N; ++i) x1[i] = in[i] + h * dxdt[i];
    sys.sys(x1, d, th);
    x0 = in;
#pragma unroll 1
    for (int s = 1; s != steps; ++s) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double tmp = x1[i];
        x1[i] = x0[i] + h2 * d[i];
        x0[i] = tmp;
      }
      th += h;
      sys.sys(x1, d, th);
    }
    const double c3 = (1.0 / 2.0) * h;
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = (1.0 / 2.0) * x0[i] + (1.0 / 2.0) * x1[i] + c3 * d[i];
  }

  ODEINTR_DEV double calc_h_opt(const double h, const double error, const int k) const {
    const double STEPFAC1 = 0.65, STEPFAC2 = 0.94, STEPFAC3 = 0.02, STEPFAC4 = 4.0;
    const double expo = 1.0 / (2 * k + 1);
    const double facmin = pow(STEPFAC3, expo);
    double fac;
    if (error == 0.0) fac = 1.0 / facmin;
    else {
      fac = STEPFAC2 / pow(error / STEPFAC1, expo);
      fac = fmax(facmin / STEPFAC4, fmin(1.0 / facmin, fac));
    }
    return h * fac;
  }
  ODEINTR_DEV bool should_reject(const double error, const int k) const {
    if (k == k_opt - 1) {
      const double d = interval(k_opt) * interval(k_opt + 1) / (interval(0) * interval(0));
      return error > d * d;
    } else if (k == k_opt) {
      const double d = interval(k_opt) / interval(0);
      return error > d * d;
    }
    return error > 1.0;
  }

  ODEINTR_DEV bool try_step(Sys& sys, vec<N>& x, double& t, double& dt) {
    const double KFAC2 = 0.9;
    vec<N> dxdt, out, err;
    sys.sys(x, dxdt, t);
    if (dt != dt_last) reset();  // step size changed from outside -> reset
    bool reject = true;
    double h_opt[K_MAX + 1], work[K_MAX + 1];
    double new_h = dt;
    bool done = false;
#pragma unroll
    for (int k = 0; k <= K_MAX; ++k) {
      if (!done && k <= k_opt + 1) {
        if (k == 0) {
          midpoint(sys, x, dxdt, t, out, dt, 2);
        } else {
          midpoint(sys, x, dxdt, t, table[k - 1], dt, 2 * (k + 1));
          // extrapolate(k, table, coeff, out): coeff[k][j] = 1 / (r*r - 1), r = n_k / n_j
#pragma unroll
          for (int j = k - 1; j > 0; --j) {
            const double r = interval(k) / interval(j);
            const double cf = 1.0 / (r * r - 1.0);
            const double a1 = 1.0 + cf, a2 = -cf;
#pragma unroll
            for (int i = 0; i < N; ++i) table[j - 1][i] = a1 * table[j][i] + a2 * table[j - 1][i];
          }
          {
            const double r = interval(k) / interval(0);
            const double cf = 1.0 / (r * r - 1.0);
            const double a1 = 1.0 + cf, a2 = -cf;
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = a1 * table[0][i] + a2 * out[i];
          }
#pragma unroll
          for (int i = 0; i < N; ++i) err[i] = out[i] + (-1.0) * table[0][i];
          const double error = reduce_error(sys, error_norm<N>(x, dxdt, err, dt, eps_abs, eps_rel), 0);
          h_opt[k] = calc_h_opt(dt, error, k);
          work[k] = cost(k) / h_opt[k];
          if ((k == k_opt - 1) || first) {  // convergence before k_opt
            if (error < 1.0) {
              reject = false;
              if ((work[k] < KFAC2 * work[k - 1]) || (k_opt <= 2)) {
                k_opt = min(K_MAX - 1, max(2, k + 1));
                new_h = h_opt[k];
                new_h *= cost(k + 1) / cost(k);
              } else {
                k_opt = min(K_MAX - 1, max(2, k));
                new_h = h_opt[k];
              }
              done = true;
            } else if (should_reject(error, k) && !first) {
              reject = true;
              new_h = h_opt[k];
              done = true;
            }
          }
          if (!done && k == k_opt) {  // convergence at k_opt
            if (error < 1.0) {
              reject = false;
              if (work[k - 1] < KFAC2 * work[k]) {
                k_opt = max(2, k_opt - 1);
                new_h = h_opt[k_opt];
              } else if ((work[k] < KFAC2 * work[k - 1]) && !last_step_rejected) {
                k_opt = min(K_MAX - 1, k_opt + 1);
                new_h = h_opt[k];
                new_h *= cost(k_opt) / cost(k);
              } else
                new_h = h_opt[k_opt];
              done = true;
            } else if (should_reject(error, k)) {
              reject = true;
              new_h = h_opt[k_opt];
              done = true;
            }
          }
          if (!done && k == k_opt + 1) {  // convergence at k_opt+1
            if (error < 1.0) {
              reject = false;
              if (work[k - 2] < KFAC2 * work[k - 1]) k_opt = max(2, k_opt - 1);
              if ((work[k] < KFAC2 * work[k_opt]) && !last_step_rejected) k_opt = min(K_MAX - 1, k);
              new_h = h_opt[k_opt];
            } else {
              reject = true;
              new_h = h_opt[k_opt];
            }
            done = true;
          }
        }
      }
    }
    if (!reject) t += dt;
    if (!last_step_rejected || less_with_sign(new_h, dt, dt)) {
      dt_last = new_h;
      dt = new_h;
    }
    last_step_rejected = reject;
    first = false;
    if (reject) { cnt.rejected++; return false; }
    x = out;
    cnt.accepted++;
    return true;
  }
};

}  // namespace odeintr_b200
