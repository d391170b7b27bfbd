// oracle/oracle_template.cpp — TEST INFRASTRUCTURE ONLY (see oracle/odeint_restated.hpp).
//
// Host translation unit the oracle builder (oracle/build_oracle.py) fills in and compiles with
// `g++ -O2 -ffp-contract=off`: the same user snippet the product compiles for the GPU, run through the
// CPU restatement of Boost odeint.  It plays the role of the module the reference generates from
// inst/templates/compile_sys_template.cpp / compile_implicit_template.cpp (which cannot be built here:
// no R, Rcpp or Boost in this image), reduced to a C interface the tests and bench.py's cpu_baseline
// leg can call through ctypes.  Only tests/, __graft_entry__.smoke() and bench.py may use it.
//
// Placeholders: @SYS_SIZE@ @METHOD@ @ATOL@ @RTOL@ @GLOBALS@ @SYS@ @JACOBIAN@ @SET_PARS@ @NPARS@
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>
#include <array>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "odeint_restated.hpp"

#define ORACLE_EULER 0
#define ORACLE_RK4 1
#define ORACLE_RK54 2
#define ORACLE_RK5 3
#define ORACLE_RK78 4
#define ORACLE_RK5_A 10
#define ORACLE_RK5_I 11
#define ORACLE_RK54_A 12
#define ORACLE_RK78_A 13
#define ORACLE_BS 14
#define ORACLE_BSD 15
#define ORACLE_ROSENBROCK4 20
#define ORACLE_METHOD @METHOD@

namespace odeintr {
using restated::state_type;
static const std::size_t N = @SYS_SIZE@;
static const double atol = @ATOL@, rtol = @RTOL@;
typedef restated::matrix sys_mat;
typedef state_type sys_vec;

// parameters are thread_local so that the OpenMP-over-instances ensemble loop can give every
// instance its own values (the reference keeps them in namespace-scope statics, one system at a time)
@GLOBALS@;

static const int M = static_cast<int>(std::sqrt(static_cast<double>(N)));
static inline int bound(int c) { int r = c % M; return r < 0 ? r + M : r; }
static inline double laplace2D(const state_type& x, int i, int j) {
  const double c = x[i * M + j];
  return (x[i * M + bound(j - 1)] - c) + (x[bound(i - 1) * M + j] - c) + (x[i * M + bound(j + 1)] - c) +
         (x[bound(i + 1) * M + j] - c);
}

// the helper the reference's documentation promises (R/odeintr.R:383-388) but its sources never define
static inline void laplace4(const state_type& x, state_type& dxdt, double D) {
  for (std::size_t i = 0; i < N; ++i) dxdt[i] = D * laplace2D(x, static_cast<int>(i) / M, static_cast<int>(i) % M);
}

struct rhs_t {
  restated::counters* cnt;
  void operator()(const state_type& x, state_type& dxdt, const double t) const {
    if (cnt) cnt->rhs++;
    @SYS@;
  }
};
struct jac_t {
  void operator()(const state_type& x, sys_mat& J, const double t, state_type& dfdt) const {
    (void)x; (void)J; (void)t; (void)dfdt;
    @JACOBIAN@;
  }
};
struct implicit_pair {
  rhs_t first;
  jac_t second;
};

static void set_pars(const double* odeintr_p, std::size_t odeintr_ld, std::size_t odeintr_i) {
  (void)odeintr_p; (void)odeintr_ld; (void)odeintr_i;
  @SET_PARS@;
}

struct recorder {
  double* t;
  double* x;         // rows x N, row-major (one trajectory)
  std::size_t cap, rows;
  void operator()(const state_type& s, double time) {
    if (rows < cap) {
      if (t) t[rows] = time;
      if (x) for (std::size_t i = 0; i < N; ++i) x[rows * N + i] = s[i];
    }
    ++rows;
  }
};

#if ORACLE_METHOD == ORACLE_EULER
typedef restated::euler stepper_t;
#define ORACLE_FAMILY_PLAIN 1
static stepper_t make_stepper(restated::counters*) { return stepper_t(); }
#elif ORACLE_METHOD == ORACLE_RK4
typedef restated::runge_kutta4 stepper_t;
#define ORACLE_FAMILY_PLAIN 1
static stepper_t make_stepper(restated::counters*) { return stepper_t(); }
#elif ORACLE_METHOD == ORACLE_RK54
typedef restated::runge_kutta_cash_karp54 stepper_t;
#define ORACLE_FAMILY_PLAIN 1
static stepper_t make_stepper(restated::counters*) { return stepper_t(); }
#elif ORACLE_METHOD == ORACLE_RK5
typedef restated::runge_kutta_dopri5 stepper_t;
#define ORACLE_FAMILY_PLAIN 1
static stepper_t make_stepper(restated::counters*) { return stepper_t(); }
#elif ORACLE_METHOD == ORACLE_RK78
typedef restated::runge_kutta_fehlberg78 stepper_t;
#define ORACLE_FAMILY_PLAIN 1
static stepper_t make_stepper(restated::counters*) { return stepper_t(); }
#elif ORACLE_METHOD == ORACLE_RK78_A
typedef restated::controlled_fehlberg78 stepper_t;
#define ORACLE_FAMILY_CONTROLLED 1
static stepper_t make_stepper(restated::counters* c) { stepper_t s(atol, rtol); s.cnt = c; return s; }
#elif ORACLE_METHOD == ORACLE_BS
typedef restated::bulirsch_stoer stepper_t;
#define ORACLE_FAMILY_CONTROLLED 1
static stepper_t make_stepper(restated::counters* c) { stepper_t s; s.cnt = c; return s; }  // default eps: atol/rtol ignored
#elif ORACLE_METHOD == ORACLE_BSD
typedef restated::bulirsch_stoer_dense_out stepper_t;
#define ORACLE_FAMILY_DENSE 1
static stepper_t make_stepper(restated::counters* c) { stepper_t s; s.cnt = c; return s; }  // default eps: atol/rtol ignored
#elif ORACLE_METHOD == ORACLE_RK5_A
typedef restated::controlled_dopri5 stepper_t;
#define ORACLE_FAMILY_CONTROLLED 1
static stepper_t make_stepper(restated::counters* c) { stepper_t s(atol, rtol); s.cnt = c; return s; }
#elif ORACLE_METHOD == ORACLE_RK54_A
typedef restated::controlled_cash_karp54 stepper_t;
#define ORACLE_FAMILY_CONTROLLED 1
static stepper_t make_stepper(restated::counters* c) { stepper_t s(atol, rtol); s.cnt = c; return s; }
#elif ORACLE_METHOD == ORACLE_RK5_I
typedef restated::dense_dopri5 stepper_t;
#define ORACLE_FAMILY_DENSE 1
static stepper_t make_stepper(restated::counters* c) { stepper_t s(atol, rtol); s.m_stepper.cnt = c; return s; }
#elif ORACLE_METHOD == ORACLE_ROSENBROCK4
typedef restated::rosenbrock4_dense_output stepper_t;
#define ORACLE_FAMILY_DENSE 1
#define ORACLE_IMPLICIT 1
static stepper_t make_stepper(restated::counters* c) { stepper_t s(atol, rtol); s.m_stepper.cnt = c; return s; }
#endif

#ifdef ORACLE_IMPLICIT
typedef implicit_pair system_t;
static system_t make_system(restated::counters* c) { system_t s; s.first.cnt = c; return s; }
#else
typedef rhs_t system_t;
static system_t make_system(restated::counters* c) { system_t s; s.cnt = c; return s; }
#endif

// The three integrate_* entry points, dispatched on the stepper family exactly as Boost's tag dispatch does.
template <class Obs>
static std::size_t run_const(state_type& x, double t0, double t1, double dt, Obs& obs, restated::counters* c) {
  stepper_t st = make_stepper(c);  // integrate_* take the stepper by value: a fresh copy per call
  system_t sys = make_system(c);
#if defined(ORACLE_FAMILY_PLAIN)
  return restated::integrate_const_plain(st, sys, x, t0, t1, dt, obs);
#elif defined(ORACLE_FAMILY_CONTROLLED)
  return restated::integrate_const_controlled(st, sys, x, t0, t1, dt, obs);
#else
  return restated::integrate_const_dense(st, sys, x, t0, t1, dt, obs);
#endif
}
template <class Obs>
static std::size_t run_times(state_type& x, const double* tb, const double* te, double dt, Obs& obs,
                             restated::counters* c) {
  stepper_t st = make_stepper(c);
  system_t sys = make_system(c);
#if defined(ORACLE_FAMILY_PLAIN)
  return restated::integrate_times_plain(st, sys, x, tb, te, dt, obs);
#elif defined(ORACLE_FAMILY_CONTROLLED)
  return restated::integrate_times_controlled(st, sys, x, tb, te, dt, obs);
#else
  return restated::integrate_times_dense(st, sys, x, tb, te, dt, obs);
#endif
}
template <class Obs>
static std::size_t run_adaptive(state_type& x, double t0, double t1, double dt, Obs& obs, restated::counters* c) {
  stepper_t st = make_stepper(c);
  system_t sys = make_system(c);
#if defined(ORACLE_FAMILY_PLAIN)
  return restated::integrate_adaptive_plain(st, sys, x, t0, t1, dt, obs);
#elif defined(ORACLE_FAMILY_CONTROLLED)
  return restated::integrate_adaptive_controlled(st, sys, x, t0, t1, dt, obs);
#else
  return restated::integrate_adaptive_dense(st, sys, x, t0, t1, dt, obs);
#endif
}

}  // namespace odeintr

extern "C" {

int oracle_sys_dim() { return (int)odeintr::N; }
int oracle_n_pars() { return @NPARS@; }
int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Single trajectory.  kind: 0 integrate_const, 1 integrate_times, 2 integrate_adaptive.
// x: N doubles in/out.  pars: NPARS doubles or null.  rec_t / rec_x: cap rows (may be null = no observer).
// counters[0..2] = accepted, rejected, rhs evaluations; returns rows observed, or -1 on step overflow.
long oracle_run(int kind, double* x, const double* pars, double t0, double t1, double dt, const double* times,
                std::size_t n_times, double* rec_t, double* rec_x, std::size_t cap, long* counters) {
  using namespace odeintr;
  if (pars) set_pars(pars, 1, 0);
  restated::counters cnt;
  state_type s(x, x + N);
  recorder rec{rec_t, rec_x, cap, 0};
  restated::null_observer nobs;
  const bool record = (rec_t != nullptr) || (rec_x != nullptr) || cap > 0;
  try {
    if (kind == 0) {
      if (record) run_const(s, t0, t1, dt, rec, &cnt); else run_const(s, t0, t1, dt, nobs, &cnt);
    } else if (kind == 1) {
      if (record) run_times(s, times, times + n_times, dt, rec, &cnt); else run_times(s, times, times + n_times, dt, nobs, &cnt);
    } else {
      if (record) run_adaptive(s, t0, t1, dt, rec, &cnt); else run_adaptive(s, t0, t1, dt, nobs, &cnt);
    }
  } catch (const restated::step_overflow&) {
    return -1;
  }
  for (std::size_t i = 0; i < N; ++i) x[i] = s[i];
  if (counters) { counters[0] = cnt.accepted; counters[1] = cnt.rejected; counters[2] = cnt.rhs; }
  return (long)rec.rows;
}

// One raw rosenbrock4::do_step with a FIXED step size (no controller): xout = 4th-order result, xerr = the embedded
// error estimate (xout - xerr is the 3rd-order companion), dense = interpolant at t + theta * dt.  Lets the tests
// measure the orders of the Ross coefficients through the very arithmetic every implicit parity test relies on.
// Returns -1 for explicit methods.
int oracle_rosenbrock4_step(const double* x, const double* pars, double t, double dt, double theta, double* xout,
                            double* xerr, double* dense) {
#ifdef ORACLE_IMPLICIT
  using namespace odeintr;
  if (pars) set_pars(pars, 1, 0);
  restated::rosenbrock4 st;
  system_t sys = make_system(nullptr);
  state_type xi(x, x + N), xo, xe, xd;
  st.do_step(sys, xi, t, xo, dt, xe);
  st.prepare_dense_output();
  st.calc_state(t + theta * dt, xd, xi, t, xo, t + dt);
  for (std::size_t i = 0; i < N; ++i) { xout[i] = xo[i]; xerr[i] = xe[i]; dense[i] = xd[i]; }
  return 0;
#else
  (void)x; (void)pars; (void)t; (void)dt; (void)theta; (void)xout; (void)xerr; (void)dense;
  return -1;
#endif
}

// Ensemble: what a user of the reference does with an R loop over instances (README.md:157-163), here
// as a C loop, optionally `#pragma omp parallel for` over instances ("all host cores" baseline).
// X: N x m SoA in/out; P: NPARS x m SoA (par_inc 1) or NPARS values (par_inc 0) or null.
// out: rows x N x m (sample-major SoA, the device layout) or null.  obs_stride: keep every k-th
// observer call of integrate_const (plus the last).  acc/rej: per-instance counters or null.
long oracle_ensemble(int kind, double* X, const double* P, std::size_t m, int par_inc, double t0, double t1,
                     double dt, const double* times, std::size_t n_times, long obs_stride, double* out,
                     std::size_t out_rows, long* acc, long* rej, int threads) {
  using namespace odeintr;
  long bad = 0;
  long rows_seen = 0;
  (void)threads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads) reduction(+ : bad) reduction(max : rows_seen)
  for (long i = 0; i < (long)m; ++i) {
    if (P) set_pars(P, par_inc ? m : 1, par_inc ? (std::size_t)i : 0);
    restated::counters cnt;
    state_type s(N);
    for (std::size_t v = 0; v < N; ++v) s[v] = X[v * m + i];
    std::size_t calls = 0, row = 0;
    auto obs = [&](const state_type& st, double) {
      if (out) {
        const bool keep = (obs_stride <= 1) || (calls % (std::size_t)obs_stride == 0);
        if (keep && row < out_rows) {
          for (std::size_t v = 0; v < N; ++v) out[(row * N + v) * m + i] = st[v];
          ++row;
        }
      }
      ++calls;
    };
    restated::null_observer nobs;
    try {
      if (kind == 0) { if (out) run_const(s, t0, t1, dt, obs, &cnt); else run_const(s, t0, t1, dt, nobs, &cnt); }
      else if (kind == 1) { if (out) run_times(s, times, times + n_times, dt, obs, &cnt); else run_times(s, times, times + n_times, dt, nobs, &cnt); }
      else { if (out) run_adaptive(s, t0, t1, dt, obs, &cnt); else run_adaptive(s, t0, t1, dt, nobs, &cnt); }
      // with a stride the final observation is always kept (it is the state at t1)
      if (out && obs_stride > 1 && calls > 0 && (calls - 1) % (std::size_t)obs_stride != 0 && row < out_rows) {
        for (std::size_t v = 0; v < N; ++v) out[(row * N + v) * m + i] = s[v];
        ++row;
      }
    } catch (const restated::step_overflow&) {
      bad += 1;
    }
    for (std::size_t v = 0; v < N; ++v) X[v * m + i] = s[v];
    if (acc) acc[i] = cnt.accepted;
    if (rej) rej[i] = cnt.rejected;
    if ((long)row > rows_seen) rows_seen = (long)row;
  }
  return bad ? -bad : rows_seen;
}

}  // extern "C"
