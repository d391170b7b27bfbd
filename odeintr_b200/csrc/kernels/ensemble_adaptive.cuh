// odeintr_b200/csrc/kernels/ensemble_adaptive.cuh
// K2: adaptive ensemble kernel — the device replacement for integrate_const / integrate_times /
// integrate_adaptive (compile_sys_template.cpp:107,121,123,134,136,154,168) when the stepper is a
// controlled or dense-output stepper: rk5_a, rk5_i (the reference's default method), rk54_a, and the
// rosenbrock4 dense-output stepper of compile_implicit.
//
// One trajectory per thread.  Step size, error estimate, accepted / rejected counters and the dense-output
// interpolant are per instance and live in registers; the only global traffic is the initial state, the
// requested samples (SoA rows, instance index fastest) and the final state + counters.
//
// Included at the end of the generated translation unit with ODEINTR_SYS_SIZE and
// ODEINTR_STEPPER (dopri5_dense | dopri5_controlled | rk54_controlled | rosenbrock4_dense).
#pragma once
#include "odeintr_b200/integrate_device.cuh"

namespace odeintr_b200 {
template <bool B, class T> struct enable_if_c {};
template <class T> struct enable_if_c<true, T> { typedef T type; };
}  // namespace odeintr_b200

// Register budget: measured on a B200 (scripts/explore_regs.py, 10 M instances) 5 resident blocks of 128 threads
// per SM (<= 96 registers, a few spilled words in L1) beats the uncapped build by 3 % (dopri5, 104 registers) to
// 12 % (rosenbrock4, 156 registers); tighter caps lose again.
// Round 2 (profiles/r02_bench_outline.log): rosenbrock4 is better off with 4 blocks of 128 registers (88 B of spills
// instead of 240: 1.02e10 -> 1.09e10 accepted steps/s on config 4); dopri5 keeps 5 (4.23e10 against 4.05e10).
#ifndef ODEINTR_ADAPTIVE_MIN_BLOCKS
#ifdef ODEINTR_IMPLICIT_SYSTEM
#define ODEINTR_ADAPTIVE_MIN_BLOCKS ((ODEINTR_MIN_BLOCKS) > 1 ? (ODEINTR_MIN_BLOCKS) : (ODEINTR_BLOCK == 128 ? 4 : 1))
#else
#define ODEINTR_ADAPTIVE_MIN_BLOCKS ((ODEINTR_MIN_BLOCKS) > 1 ? (ODEINTR_MIN_BLOCKS) : (ODEINTR_BLOCK == 128 ? 5 : 1))
#endif
#endif

extern "C" __global__ void __launch_bounds__(ODEINTR_BLOCK, ODEINTR_ADAPTIVE_MIN_BLOCKS)
odeintr_adaptive_ensemble(const odeintr_adaptive_args a) {
  constexpr int N = ODEINTR_SYS_SIZE;
  typedef odeintr::system_type Sys;
  typedef odeintr_b200::ODEINTR_STEPPER<Sys, N> Stepper;
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= a.m) return;
  const unsigned live = __activemask();  // lanes of this warp that own an instance
  // K6: with a cost-proxy order (odeintr_set_instance_order) neighbouring lanes hold trajectories that step alike
  const long long i = a.perm ? static_cast<long long>(a.perm[slot]) : slot;

  Sys sys;
  sys.load_params(a.pars, a.par_ld, i * a.par_inc);
  Stepper st;
  st.setup(a.atol, a.rtol);
  st.counters().budget = a.max_tries;

  odeintr_b200::vec<N> x;
#pragma unroll
  for (int v = 0; v < N; ++v) x[v] = a.x[v * a.ld + i];

  odeintr_b200::sample_writer<N> obs;
  obs.out = a.out ? a.out + i : nullptr;
  obs.t_rec = a.t_rec ? a.t_rec + i : nullptr;
  obs.ld = a.ld;
  obs.rows = 0;
  obs.max_rows = a.max_rows;

  bool ok;
  if constexpr (Stepper::is_dense) {
#ifdef ODEINTR_NESTED_LOOPS  // Boost's literal loop nesting (kept for A/B measurements)
    if (a.mode == ODEINTR_MODE_TIMES) ok = odeintr_b200::integrate_times_dense(st, sys, x, a.times, a.n_times, a.dt, obs);
    else if (a.mode == ODEINTR_MODE_CONST) ok = odeintr_b200::integrate_const_dense(st, sys, x, a.t0, a.t1, a.dt, obs);
    else ok = odeintr_b200::integrate_adaptive_dense(st, sys, x, a.t0, a.t1, a.dt, obs);
#else
    if (a.mode == ODEINTR_MODE_TIMES) ok = odeintr_b200::flat_times_dense(st, sys, x, a.times, a.n_times, a.dt, obs);
    else if (a.mode == ODEINTR_MODE_CONST) ok = odeintr_b200::flat_const_dense(st, sys, x, a.t0, a.t1, a.dt, obs);
    else ok = odeintr_b200::flat_adaptive_dense(st, sys, x, a.t0, a.t1, a.dt, obs);
#endif
  } else {
    double dt = a.dt;
    if (a.mode == ODEINTR_MODE_TIMES) ok = odeintr_b200::integrate_times_controlled(st, sys, x, a.times, a.n_times, dt, obs);
    else if (a.mode == ODEINTR_MODE_CONST) ok = odeintr_b200::integrate_const_controlled(st, sys, x, a.t0, a.t1, dt, obs);
    else {
      double t = a.t0;
      ok = odeintr_b200::integrate_adaptive_controlled(st, sys, x, t, a.t1, dt, obs);
    }
  }
  (void)ok;  // the status word carries the failure
#pragma unroll
  for (int v = 0; v < N; ++v) a.x[v * a.ld + i] = x[v];
  odeintr_b200::step_counters& c = st.counters();
  if (a.accepted) a.accepted[i] = c.accepted;
  if (a.rejected) a.rejected[i] = c.rejected;
  if (a.status) {
    int status = c.status;
#pragma unroll
    for (int v = 0; v < N; ++v) if (!isfinite(x[v]) && status == ODEINTR_STATUS_OK) status = ODEINTR_STATUS_NONFINITE;
    a.status[i] = status;
    if (status != ODEINTR_STATUS_OK && a.fail_flags) atomicOr(a.fail_flags, status);
  }
  if (a.rows) a.rows[i] = obs.rows;
  if (a.rows_min) {
    // integrate_const with a dense-output stepper may or may not emit the sample at the very end of the interval
    // (it depends on whether the last real step was taken while that sample was still ahead): the host keeps the
    // rows every instance produced
    __syncwarp(live);
    const int least = __reduce_min_sync(live, obs.rows);
    if ((threadIdx.x & 31) == (__ffs(live) - 1)) atomicMin(a.rows_min, least);
  }
}

namespace odeintr_b200 {
// step size after `attempts` step attempts from (x, t0, dt0): dense-output steppers ...
template <class St, class Sys, int N>
ODEINTR_DEV typename enable_if_c<St::is_dense, double>::type pilot_step_size(St& st, Sys& sys, vec<N>& x, const double t0,
                                                                            const double dt0, const int attempts) {
  st.initialize(x, t0, dt0);
  bool in_step = false;
  for (int k = 0; k < attempts; ++k) {
    if (!in_step) { st.begin_step(sys); in_step = true; }
    const int r = st.try_once(sys);
    if (r > 0) in_step = false;
    else if (r < 0) break;
  }
  return st.dt;
}
// ... and controlled steppers
template <class St, class Sys, int N>
ODEINTR_DEV typename enable_if_c<!St::is_dense, double>::type pilot_step_size(St& st, Sys& sys, vec<N>& x, const double t0,
                                                                             const double dt0, const int attempts) {
  double t = t0, dt = dt0;
  for (int k = 0; k < attempts; ++k) {
    if (st.counters().exhausted()) break;
    st.try_step(sys, x, t, dt);
  }
  return dt;
}
}  // namespace odeintr_b200

// K6 pilot: the first `attempts` step attempts of every instance on a scratch copy of its state; key[i] = the step
// size the controller has settled on by then (bit pattern of |dt| as a float: positive floats order like unsigned
// integers).  Trajectories with the same key advance at the same rate -- they reject, accept and cross their sample
// times together -- so sorting the instances by it (odeintr_set_instance_order) restores the lane coherence of a sorted
// parameter sweep for ensembles that arrive shuffled.  Nothing is written back but the key.
extern "C" __global__ void __launch_bounds__(ODEINTR_BLOCK, ODEINTR_ADAPTIVE_MIN_BLOCKS)
odeintr_adaptive_pilot(const odeintr_adaptive_args a, unsigned* key, const int attempts) {
  constexpr int N = ODEINTR_SYS_SIZE;
  typedef odeintr::system_type Sys;
  typedef odeintr_b200::ODEINTR_STEPPER<Sys, N> Stepper;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.m) return;
  Sys sys;
  sys.load_params(a.pars, a.par_ld, i * a.par_inc);
  Stepper st;
  st.setup(a.atol, a.rtol);
  st.counters().budget = a.max_tries;
  odeintr_b200::vec<N> x;
#pragma unroll
  for (int v = 0; v < N; ++v) x[v] = a.x[v * a.ld + i];
  const double dt = odeintr_b200::pilot_step_size(st, sys, x, a.t0, a.dt, attempts);
  const float mag = fabsf(static_cast<float>(dt));
  // near-equal step sizes tie (the low mantissa bits are dropped) and a stable sort keeps their original order
  key[i] = (mag == mag) ? (__float_as_uint(mag) >> 10) : 0xffffffffu;
}
