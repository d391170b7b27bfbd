// odeintr_b200/csrc/kernels/bulirsch_stoer_dense.cuh
// Bulirsch-Stoer with dense output, one trajectory per thread: Boost's bulirsch_stoer_dense_out<state_type> over
// modified_midpoint_dense_out (method "bsd", R/utils.R:58; the stepper of the reference README's Van der Pol
// demos, README.md:125,155).  Default-constructed by the reference (eps 1e-6 / 1e-6 whatever atol / rtol say, no
// interpolation control).  Interval sequence 2, 6, 10, ...; all derivatives of every midpoint run are kept
// (162 vectors), their central finite differences at the interval centre are extrapolated like the states, and
// the dense output is the Taylor polynomial about the centre.  That working set (~350 N doubles per trajectory)
// cannot live in registers: it sits in thread-local memory (L1/L2-backed) and is addressed dynamically, so this
// stepper is functional rather than fast; the dopri5 and rosenbrock4 dense steppers are the tuned ones.
#pragma once
#include "odeintr_b200/adaptive_steppers.cuh"

namespace odeintr_b200 {

template <class Sys, int N>
struct bsd_dense {
  static constexpr bool is_dense = true;
  static constexpr int K_MAX = 8;
  double eps_abs, eps_rel;
  step_counters cnt;
  bool last_step_rejected, first;
  int k_opt, k_final, fails;
  double t, dt, t_old;
  vec<N> x_cur, x_old, d_cur, d_old;
  vec<N> table[K_MAX];
  vec<N> mp[K_MAX + 1];
  vec<N> derivs[2 * (K_MAX + 1) * (K_MAX + 1)];  // run k starts at 2 k^2 and holds 2 + 4k derivatives
  vec<N> diffs[2 * K_MAX + 2][K_MAX + 1];

  ODEINTR_DEV void setup(const double, const double) {
    eps_abs = 1e-6; eps_rel = 1e-6;  // the reference's stepper_type() ignores atol / rtol for bsd
    cnt.reset();
    k_opt = 4; k_final = 0;
    first = true; last_step_rejected = false;
  }
  ODEINTR_DEV step_counters& counters() { return cnt; }
  ODEINTR_DEV void initialize(const vec<N>& x0, const double t0, const double dt0) {
    x_cur = x0; t = t0; dt = dt0;
    first = true; last_step_rejected = false;  // reset()
  }
  ODEINTR_DEV const vec<N>& current_state() const { return x_cur; }

  static ODEINTR_DEV int steps_of(const int k) { return 2 + 4 * k; }
  static ODEINTR_DEV double cost(const int k) { return static_cast<double>(2 * (k + 1) * (k + 1)); }  // sum of 2 + 4j
  static ODEINTR_DEV double coeff(const int i, const int k) {
    const double r = static_cast<double>(steps_of(i)) / static_cast<double>(steps_of(k));
    return 1.0 / (r * r - 1.0);
  }
  ODEINTR_DEV vec<N>& deriv(const int k, const int i) { return derivs[2 * k * k + i]; }

  // modified_midpoint_dense_out::do_step
  ODEINTR_DEV void midpoint(Sys& sys, const int k, const vec<N>& in, const vec<N>& dxdt, const double tt, vec<N>& out,
                            const double h_total) {
    const int steps = steps_of(k);
    const double h = h_total / static_cast<double>(steps);
    const double h2 = 2.0 * h;
    double th = tt + h;
    vec<N> x0, x1;
#pragma unroll
    for (int q = 0; q < N; ++q) x1[q] = in[q] + h * dxdt[q];
    if (steps == 2) mp[k] = x1;
    sys.sys(x1, deriv(k, 0), th);
    x0 = in;
#pragma unroll 1
    for (int i = 1; i != steps; ++i) {
      const vec<N>& d = deriv(k, i - 1);
#pragma unroll
      for (int q = 0; q < N; ++q) {
        const double tmp = x1[q];
        x1[q] = x0[q] + h2 * d[q];
        x0[q] = tmp;
      }
      if (i == steps / 2 - 1) mp[k] = x1;
      th += h;
      sys.sys(x1, deriv(k, i), th);
    }
    const vec<N>& dl = deriv(k, steps - 1);
    const double c3 = (1.0 / 2.0) * h;
#pragma unroll
    for (int q = 0; q < N; ++q) out[q] = (1.0 / 2.0) * x0[q] + (1.0 / 2.0) * x1[q] + c3 * dl[q];
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
      const double d = static_cast<double>(steps_of(k_opt) * steps_of(k_opt + 1) / (steps_of(0) * steps_of(0)));
      return error > d * d;
    } else if (k == k_opt) {
      const double d = static_cast<double>(steps_of(k_opt) / steps_of(0));
      return error > d * d;
    }
    return error > 1.0;
  }
  static ODEINTR_DEV double binomial(const int n, const int k) {
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * static_cast<double>(n - k + i) / static_cast<double>(i);
    return floor(r + 0.5);
  }
  ODEINTR_DEV void finite_difference(const int j, const int kappa, const double fac, const vec<N>& dxdt) {
    const int m = steps_of(j) / 2 - 1;
    if (kappa == 0) {
      const vec<N>& s = deriv(j, m);
#pragma unroll
      for (int q = 0; q < N; ++q) diffs[0][j][q] = fac * s[q];
    } else {
      vec<N>& d = diffs[kappa][j - kappa / 2];
      const vec<N>& s = deriv(j, m + kappa);
#pragma unroll
      for (int q = 0; q < N; ++q) d[q] = fac * s[q];
      double sign = -1.0;
      int c = 1;
#pragma unroll 1
      for (int i = m + kappa - 2; i >= m - kappa; i -= 2) {
        if (i >= 0) {
          const double w = sign * fac * binomial(kappa, c);
          const vec<N>& si = deriv(j, i);
#pragma unroll
          for (int q = 0; q < N; ++q) d[q] = d[q] + w * si[q];
        } else {
          const double w = sign * fac;
#pragma unroll
          for (int q = 0; q < N; ++q) d[q] = d[q] + w * dxdt[q];
        }
        sign *= -1;
        ++c;
      }
    }
  }
  // extrapolate_dense_out(k, table, coeff, order_start_index) on a table of vec<N>
  ODEINTR_DEV void extrapolate_table(const int k, vec<N>* tab, const int start) {
#pragma unroll 1
    for (int j = k; j > 1; --j) {
      const double cf = coeff(k + start, j + start - 1);
      const double a1 = 1.0 + cf, a2 = -cf;
#pragma unroll
      for (int q = 0; q < N; ++q) tab[j - 1][q] = a1 * tab[j][q] + a2 * tab[j - 1][q];
    }
    const double cf = coeff(k + start, start);
    const double a1 = 1.0 + cf, a2 = -cf;
#pragma unroll
    for (int q = 0; q < N; ++q) tab[0][q] = a1 * tab[1][q] + a2 * tab[0][q];
  }
  ODEINTR_DEV void prepare_dense_output(const int k, const vec<N>& dxdt_start, const double h) {
#pragma unroll 1
    for (int j = 0; j <= k; ++j) {
      const double d = static_cast<double>(steps_of(j)) / (2.0 * h);
      double f = 1.0;
#pragma unroll 1
      for (int kappa = 0; kappa <= 2 * j + 1; ++kappa) {
        finite_difference(j, kappa, f, dxdt_start);
        f *= d;
      }
      if (j > 0) extrapolate_table(j, mp, 0);
    }
    double d = h / 2;
#pragma unroll 1
    for (int kappa = 0; kappa <= 2 * k + 1; ++kappa) {
#pragma unroll 1
      for (int j = 1; j <= (k - kappa / 2); ++j) extrapolate_table(j, diffs[kappa], kappa / 2);
#pragma unroll
      for (int q = 0; q < N; ++q) diffs[kappa][0][q] *= d;
      d *= h / (2 * (kappa + 2));
    }
  }

  // bulirsch_stoer_dense_out::try_step(sys, in, dxdt, t, out, dxdt_new, dt)
  ODEINTR_DEV bool try_step(Sys& sys, const vec<N>& in, const vec<N>& dxdt, double& tt, vec<N>& out, vec<N>& dxdt_new,
                            double& h) {
    const double KFAC2 = 0.9;
    bool reject = true;
    double h_opt[K_MAX + 1], work[K_MAX + 1];
    k_final = 0;
    double new_h = h;
    vec<N> err;
#pragma unroll 1
    for (int k = 0; k <= k_opt + 1; ++k) {
      if (k == 0) {
        midpoint(sys, k, in, dxdt, tt, out, h);
        continue;
      }
      midpoint(sys, k, in, dxdt, tt, table[k - 1], h);
      // extrapolate(k, table, coeff, out)
#pragma unroll 1
      for (int j = k - 1; j > 0; --j) {
        const double cf = coeff(k, j);
        const double a1 = 1.0 + cf, a2 = -cf;
#pragma unroll
        for (int q = 0; q < N; ++q) table[j - 1][q] = a1 * table[j][q] + a2 * table[j - 1][q];
      }
      {
        const double cf = coeff(k, 0);
        const double a1 = 1.0 + cf, a2 = -cf;
#pragma unroll
        for (int q = 0; q < N; ++q) out[q] = a1 * table[0][q] + a2 * out[q];
      }
#pragma unroll
      for (int q = 0; q < N; ++q) err[q] = out[q] + (-1.0) * table[0][q];
      const double error = reduce_error(sys, error_norm<N>(in, dxdt, err, h, eps_abs, eps_rel), 0);
      h_opt[k] = calc_h_opt(h, error, k);
      work[k] = cost(k) / h_opt[k];
      k_final = k;
      if ((k == k_opt - 1) || first) {  // convergence before k_opt
        if (error < 1.0) {
          reject = false;
          if ((work[k] < KFAC2 * work[k - 1]) || (k_opt <= 2)) {
            k_opt = min(K_MAX - 1, max(2, k + 1));
            new_h = h_opt[k] * cost(k + 1) / cost(k);
          } else {
            k_opt = min(K_MAX - 1, max(2, k));
            new_h = h_opt[k];
          }
          break;
        } else if (should_reject(error, k) && !first) {
          reject = true;
          new_h = h_opt[k];
          break;
        }
      }
      if (k == k_opt) {  // convergence at k_opt
        if (error < 1.0) {
          reject = false;
          if (work[k - 1] < KFAC2 * work[k]) {
            k_opt = max(2, k_opt - 1);
            new_h = h_opt[k_opt];
          } else if ((work[k] < KFAC2 * work[k - 1]) && !last_step_rejected) {
            k_opt = min(K_MAX - 1, k_opt + 1);
            new_h = h_opt[k] * cost(k_opt) / cost(k);
          } else
            new_h = h_opt[k_opt];
          break;
        } else if (should_reject(error, k)) {
          reject = true;
          new_h = h_opt[k_opt];
          break;
        }
      }
      if (k == k_opt + 1) {  // convergence at k_opt+1
        if (error < 1.0) {
          reject = false;
          if (work[k - 2] < KFAC2 * work[k - 1]) k_opt = max(2, k_opt - 1);
          if ((work[k] < KFAC2 * work[k_opt]) && !last_step_rejected) k_opt = min(K_MAX - 1, k);
          new_h = h_opt[k_opt];
        } else {
          reject = true;
          new_h = h_opt[k_opt];
        }
        break;
      }
    }
    if (!reject) {
      sys.sys(out, dxdt_new, tt + h);  // dxdt for the next step and the dense output
      prepare_dense_output(k_final, dxdt, h);  // interpolation control is off in the reference: error = 0
      tt += h;
    }
    if (!last_step_rejected || (new_h < h)) h = new_h;
    last_step_rejected = reject;
    if (reject) cnt.rejected++; else cnt.accepted++;
    return !reject;
  }

  ODEINTR_DEV void begin_step(Sys& sys) {
    if (first) sys.sys(x_cur, d_cur, t);
    t_old = t;
    fails = 0;
  }
  ODEINTR_DEV int try_once(Sys& sys) {
    if (cnt.exhausted()) return -1;
    const bool ok = try_step(sys, x_cur, d_cur, t, x_old, d_old, dt);
    first = false;
    if (ok) {  // toggle_current_state()
      const vec<N> xs = x_cur, ds = d_cur;
      x_cur = x_old; d_cur = d_old;
      x_old = xs; d_old = ds;
      return 1;
    }
    if (++fails >= ODEINTR_MAX_FAILED_STEPS) { cnt.status |= ODEINTR_STATUS_STEP_OVERFLOW; return -1; }
    return 0;
  }
  ODEINTR_DEV bool do_step(Sys& sys) {
    begin_step(sys);
    int r;
    while ((r = try_once(sys)) == 0) {}
    return r > 0;
  }
  ODEINTR_DEV void calc_state(const double tq, vec<N>& out) const {
    const double theta = 2 * ((tq - t_old) / (t - t_old)) - 1;
    out = mp[0];
    double theta_pow = theta;
#pragma unroll 1
    for (int i = 0; i <= 2 * k_final + 1; ++i) {
#pragma unroll
      for (int q = 0; q < N; ++q) out[q] = out[q] + theta_pow * diffs[i][0][q];
      theta_pow *= theta;
    }
  }
};

}  // namespace odeintr_b200
