// odeintr_b200/csrc/kernels/rosenbrock4.cuh
// K3: Rosenbrock4 (Shampine / NR3 "Ross" coefficients) with step-size controller and dense output, one
// trajectory per thread.  Replaces rosenbrock4<double>, rosenbrock4_controller<> and
// rosenbrock4_dense_output<> + uBLAS lu_factorize / lu_substitute, which the reference instantiates in
// inst/templates/compile_implicit_template.cpp:28-33.  Algorithm and operation order: SURVEY.md A.6-A.8.
//
// The N x N matrix I/(gamma dt) - J, its LU factors and the pivot vector stay in registers: every loop is
// fully unrolled and row exchanges are compare-and-swap chains over compile-time indices, so no element is
// ever addressed dynamically (no local memory).  Intended for the small stiff systems compile_implicit is
// used for (N <= 8).
#pragma once
#include "odeintr_b200/adaptive_steppers.cuh"

namespace odeintr_b200 {

template <int N>
struct lu_factors {
  mat<N> m;
  int pm[N];

  // uBLAS lu_factorize with partial pivoting: first row of maximal |m(r,i)|, rows scaled by 1/pivot
  ODEINTR_DEV void factorize() {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      int piv = i;
      double best = fabs(m(i, i));
#pragma unroll
      for (int r = i + 1; r < N; ++r) {
        const double v = fabs(m(r, i));
        if (v > best) { best = v; piv = r; }
      }
      pm[i] = piv;
#pragma unroll
      for (int r = i + 1; r < N; ++r) {
        const bool sw = (piv == r);  // branch-free exchange keeps every element in a register
#pragma unroll
        for (int c = 0; c < N; ++c) {
          const double a = m(i, c), b = m(r, c);
          m(i, c) = sw ? b : a;
          m(r, c) = sw ? a : b;
        }
      }
      if (m(i, i) != 0.0) {
        const double inv = 1.0 / m(i, i);
#pragma unroll
        for (int r = i + 1; r < N; ++r) m(r, i) *= inv;
      }
#pragma unroll
      for (int r = i + 1; r < N; ++r)
#pragma unroll
        for (int c = i + 1; c < N; ++c) m(r, c) -= m(r, i) * m(i, c);
    }
  }

#ifdef ODEINTR_RECIPROCAL
  // fast mode (ODEINTR_FLAG_RECIPROCAL): the six solves of an attempt multiply by the reciprocals of U's diagonal
  double inv_diag[N];
  ODEINTR_DEV void prepare() {
#pragma unroll
    for (int n = 0; n < N; ++n) inv_diag[n] = 1.0 / m(n, n);
  }
#else
  ODEINTR_DEV void prepare() {}
#endif

  // uBLAS lu_substitute: row swaps, unit-lower forward solve, upper backward solve (column oriented)
  ODEINTR_DEV void substitute(vec<N>& v) const {
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int r = i + 1; r < N; ++r) {
        const bool sw = (pm[i] == r);
        const double a = v[i], b = v[r];
        v[i] = sw ? b : a;
        v[r] = sw ? a : b;
      }
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const double t = v[n];
#pragma unroll
      for (int r = n + 1; r < N; ++r) v[r] -= t * m(r, n);
    }
#pragma unroll
    for (int n = N - 1; n >= 0; --n) {
#ifdef ODEINTR_RECIPROCAL
      const double t = v[n] *= inv_diag[n];
#else
      // (a corrected-reciprocal quotient, exact_div.cuh, was measured here in round 2: bit-identical but no faster than
      //  the device's division sequence, and it costs registers -- 0.94e10 against 1.02e10 accepted steps/s)
      const double t = v[n] /= m(n, n);
#endif
#pragma unroll
      for (int r = 0; r < n; ++r) v[r] -= t * m(r, n);
    }
  }
};

// default_rosenbrock_coefficients<double>
struct ross_coef {
  static constexpr double gamma = 0.25;
  static constexpr double d1 = 0.25, d2 = -0.1043, d3 = 0.1035, d4 = -0.3620000000000023e-01;
  static constexpr double c2 = 0.386, c3 = 0.21, c4 = 0.63;
  static constexpr double c21 = -0.5668800000000000e+01, a21 = 0.1544000000000000e+01;
  static constexpr double c31 = -0.2430093356833875e+01, c32 = -0.2063599157091915e+00;
  static constexpr double a31 = 0.9466785280815826e+00, a32 = 0.2557011698983284e+00;
  static constexpr double c41 = -0.1073529058151375e+00, c42 = -0.9594562251023355e+01, c43 = -0.2047028614809616e+02;
  static constexpr double a41 = 0.3314825187068521e+01, a42 = 0.2896124015972201e+01, a43 = 0.9986419139977817e+00;
  static constexpr double c51 = 0.7496443313967647e+01, c52 = -0.1024680431464352e+02, c53 = -0.3399990352819905e+02,
                          c54 = 0.1170890893206160e+02;
  static constexpr double a51 = 0.1221224509226641e+01, a52 = 0.6019134481288629e+01, a53 = 0.1253708332932087e+02,
                          a54 = -0.6878860361058950e+00;
  static constexpr double c61 = 0.8083246795921522e+01, c62 = -0.7981132988064893e+01, c63 = -0.3152159432874371e+02,
                          c64 = 0.1631930543123136e+02, c65 = -0.6058818238834054e+01;
  static constexpr double d21 = 0.1012623508344586e+02, d22 = -0.7487995877610167e+01, d23 = -0.3480091861555747e+02,
                          d24 = -0.7992771707568823e+01, d25 = 0.1025137723295662e+01;
  static constexpr double d31 = -0.6762803392801253e+00, d32 = 0.6087714651680015e+01, d33 = 0.1643084320892478e+02,
                          d34 = 0.2476722511418386e+02, d35 = -0.6594389125716872e+01;
};

// dense-output Rosenbrock4 with its controller (the only stepper compile_implicit instantiates)
template <class Sys, int N>
struct rosenbrock4_dense {
  static constexpr bool is_dense = true;
  typedef ross_coef K;
  vec<N> x_cur, x_old, g1, g2, g3, g4, g5, cont3, cont4;
  double t, t_old, dt;
  double atol, rtol, err_old, dt_old;
  bool first_step, last_rejected;
  step_counters cnt;

  ODEINTR_DEV void setup(const double atol_, const double rtol_) {
    atol = atol_; rtol = rtol_;
    first_step = true; last_rejected = false;
    err_old = 0.0; dt_old = 0.0;
    cnt.reset();
  }
  ODEINTR_DEV step_counters& counters() { return cnt; }
  ODEINTR_DEV void initialize(const vec<N>& x0, const double t0, const double dt0) { x_cur = x0; t = t0; dt = dt0; }
  ODEINTR_DEV const vec<N>& current_state() const { return x_cur; }

  // rosenbrock4::do_step(sys, x, t, xout, dt, xerr)
  // (the kernel instantiates the const / times / adaptive loops, each with its own inlined copy of the attempt: 157 KB of
  //  code, and ncu shows no_instruction among the top stalls.  Measured, profiles/r02_bench_outline.log: ONE out-of-line
  //  copy halves the code but passes the vectors through local memory: 1.02e10 -> 0.96e10 accepted steps/s; kept inline.
  //  What did help is one resident block fewer: 4 blocks of 128 registers, 88 B of spills instead of 240: 1.09e10.)
  ODEINTR_DEV void stage_step(Sys& sys, const vec<N>& x, const double tt, vec<N>& xout, const double h, vec<N>& xerr) {
    vec<N> dxdt, dfdt, dn, xt;
    lu_factors<N> lu;
    sys.sys(x, dxdt, tt);
    sys.jacobian(x, lu.m, tt, dfdt);
    const double diag = 1.0 / K::gamma / h;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int c = 0; c < N; ++c) {
        lu.m(r, c) *= -1.0;
        if (r == c) lu.m(r, c) += diag * 1.0;
      }
    lu.factorize();
    lu.prepare();
#ifdef ODEINTR_RECIPROCAL
    const double rh = 1.0 / h;
#define ODEINTR_OVER_H(e) ((e) * rh)
#else
#define ODEINTR_OVER_H(e) ((e) / h)
#endif
#pragma unroll
    for (int i = 0; i < N; ++i) g1[i] = dxdt[i] + h * K::d1 * dfdt[i];
    lu.substitute(g1);
#pragma unroll
    for (int i = 0; i < N; ++i) xt[i] = x[i] + K::a21 * g1[i];
    sys.sys(xt, dn, tt + K::c2 * h);
#pragma unroll
    for (int i = 0; i < N; ++i) g2[i] = dn[i] + h * K::d2 * dfdt[i] + ODEINTR_OVER_H(K::c21 * g1[i]);
    lu.substitute(g2);
#pragma unroll
    for (int i = 0; i < N; ++i) xt[i] = x[i] + K::a31 * g1[i] + K::a32 * g2[i];
    sys.sys(xt, dn, tt + K::c3 * h);
#pragma unroll
    for (int i = 0; i < N; ++i) g3[i] = dn[i] + h * K::d3 * dfdt[i] + ODEINTR_OVER_H(K::c31 * g1[i] + K::c32 * g2[i]);
    lu.substitute(g3);
#pragma unroll
    for (int i = 0; i < N; ++i) xt[i] = x[i] + K::a41 * g1[i] + K::a42 * g2[i] + K::a43 * g3[i];
    sys.sys(xt, dn, tt + K::c4 * h);
#pragma unroll
    for (int i = 0; i < N; ++i)
      g4[i] = dn[i] + h * K::d4 * dfdt[i] + ODEINTR_OVER_H(K::c41 * g1[i] + K::c42 * g2[i] + K::c43 * g3[i]);
    lu.substitute(g4);
#pragma unroll
    for (int i = 0; i < N; ++i) xt[i] = x[i] + K::a51 * g1[i] + K::a52 * g2[i] + K::a53 * g3[i] + K::a54 * g4[i];
    sys.sys(xt, dn, tt + h);
#pragma unroll
    for (int i = 0; i < N; ++i) g5[i] = dn[i] + ODEINTR_OVER_H(K::c51 * g1[i] + K::c52 * g2[i] + K::c53 * g3[i] + K::c54 * g4[i]);
    lu.substitute(g5);
#pragma unroll
    for (int i = 0; i < N; ++i) xt[i] += g5[i];
    sys.sys(xt, dn, tt + h);
#pragma unroll
    for (int i = 0; i < N; ++i)
      xerr[i] = dn[i] + ODEINTR_OVER_H(K::c61 * g1[i] + K::c62 * g2[i] + K::c63 * g3[i] + K::c64 * g4[i] + K::c65 * g5[i]);
    lu.substitute(xerr);
#pragma unroll
    for (int i = 0; i < N; ++i) xout[i] = xt[i] + xerr[i];
#undef ODEINTR_OVER_H
  }

  // rosenbrock4_controller::try_step(sys, x, t, xout, dt)
  ODEINTR_DEV bool try_step(Sys& sys, const vec<N>& x, double& tt, vec<N>& xout, double& h) {
    const double safe = 0.9, fac1 = 5.0, fac2 = 1.0 / 6.0;
    vec<N> xerr;
    stage_step(sys, x, tt, xout, h, xerr);
    double err = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double sk = atol + rtol * fmax(fabs(x[i]), fabs(xout[i]));
#ifdef ODEINTR_RECIPROCAL
      const double e = xerr[i] * (1.0 / sk);
      err += e * e;
#else
      err += xerr[i] * xerr[i] / sk / sk;
#endif
    }
    err = sqrt(err / static_cast<double>(N));
    double fac = fmax(fac2, fmin(fac1, pow(err, 0.25) / safe));
    double h_new = h / fac;
    if (err <= 1.0) {
      if (first_step) {
        first_step = false;
      } else {
        double fac_pred = (dt_old / h) * pow(err * err / err_old, 0.25) / safe;
        fac_pred = fmax(fac2, fmin(fac1, fac_pred));
        fac = fmax(fac, fac_pred);
        h_new = h / fac;
      }
      dt_old = h;
      err_old = fmax(0.01, err);
      if (last_rejected) h_new = (h >= 0.0 ? fmin(h_new, h) : fmax(h_new, h));
      tt += h;
      h = h_new;
      last_rejected = false;
      cnt.accepted++;
      return true;
    }
    h = h_new;
    last_rejected = true;
    cnt.rejected++;
    return false;
  }

  // rosenbrock4_dense_output::do_step
  ODEINTR_DEV bool do_step(Sys& sys) {
    t_old = t;
    vec<N> xn;
    int fails = 0;
    if (cnt.exhausted()) return false;
    while (!try_step(sys, x_cur, t, xn, dt)) {
      if (++fails >= ODEINTR_MAX_FAILED_STEPS) { cnt.status |= ODEINTR_STATUS_STEP_OVERFLOW; return false; }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {  // prepare_dense_output
      cont3[i] = K::d21 * g1[i] + K::d22 * g2[i] + K::d23 * g3[i] + K::d24 * g4[i] + K::d25 * g5[i];
      cont4[i] = K::d31 * g1[i] + K::d32 * g2[i] + K::d33 * g3[i] + K::d34 * g4[i] + K::d35 * g5[i];
    }
    x_old = x_cur;
    x_cur = xn;
    return true;
  }
  // do_step split into "start a step" + "one attempt" (see dopri5_dense::try_once)
  int fails;
  ODEINTR_DEV void begin_step(Sys&) { t_old = t; fails = 0; }
  ODEINTR_DEV int try_once(Sys& sys) {
    if (cnt.exhausted()) return -1;
    vec<N> xn;
    if (try_step(sys, x_cur, t, xn, dt)) {
#pragma unroll
      for (int i = 0; i < N; ++i) {  // prepare_dense_output
        cont3[i] = K::d21 * g1[i] + K::d22 * g2[i] + K::d23 * g3[i] + K::d24 * g4[i] + K::d25 * g5[i];
        cont4[i] = K::d31 * g1[i] + K::d32 * g2[i] + K::d33 * g3[i] + K::d34 * g4[i] + K::d35 * g5[i];
      }
      x_old = x_cur;
      x_cur = xn;
      return 1;
    }
    if (++fails >= ODEINTR_MAX_FAILED_STEPS) { cnt.status |= ODEINTR_STATUS_STEP_OVERFLOW; return -1; }
    return 0;
  }

  ODEINTR_DEV void calc_state(const double tq, vec<N>& x) const {
    const double h = t - t_old;
    const double s = (tq - t_old) / h;
    const double s1 = 1.0 - s;
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = x_old[i] * s1 + s * (x_cur[i] + s1 * (cont3[i] + s * cont4[i]));
  }
};

}  // namespace odeintr_b200
