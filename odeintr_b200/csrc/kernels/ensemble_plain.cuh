// odeintr_b200/csrc/kernels/ensemble_plain.cuh
// K1: fixed-step ensemble kernel — the device replacement for
//   odeint::integrate_const(stepper, sys, state, t0, t1, dt, obs)     compile_sys_template.cpp:154
//   odeint::integrate_times(stepper, sys, state, times..., dt, obs)   compile_sys_template.cpp:123
//   odeint::integrate_adaptive(stepper, ...) for plain steppers       compile_sys_template.cpp:107,168
// when `stepper` is a plain (non-controlled) stepper: euler, rk4, rk54, rk5.
//
// One trajectory per thread; the state, the stage derivatives and the per-instance parameters live in
// registers for the whole launch.  Global traffic is the initial state (N doubles), one coalesced row
// per variable per observer sample, and the final state.  Observer rows are written with streaming
// stores: consecutive threads hold consecutive instances, so every warp store is one fully used
// 256-byte segment of a structure-of-arrays row (no shared-memory transpose needed in this layout).
//
// Included at the end of the generated translation unit, after `odeintr::system_type`, with
//   ODEINTR_SYS_SIZE, ODEINTR_STEPPER (euler_stepper | rk4_stepper | rk54_stepper | rk5_stepper | rk78_stepper)
#pragma once
#include "odeintr_b200/steppers_plain.cuh"
#include "odeintr_b200/dopri5.cuh"
#include "odeintr_b200/rk78.cuh"

extern "C" __global__ void __launch_bounds__(ODEINTR_BLOCK, ODEINTR_MIN_BLOCKS)
odeintr_plain_ensemble(const odeintr_plain_args a) {
  constexpr int N = ODEINTR_SYS_SIZE;
  typedef odeintr::system_type Sys;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.m) return;

  Sys sys;
  sys.load_params(a.pars, a.par_ld, i * a.par_inc);
  odeintr_b200::ODEINTR_STEPPER<Sys, N> stepper;
  stepper.reset();

  odeintr_b200::vec<N> x;
#pragma unroll
  for (int v = 0; v < N; ++v) x[v] = a.x[v * a.ld + i];

  const double dt = a.dt;
  double t = a.t0;
  long long step = 0;
  double* out = a.out ? a.out + i : nullptr;
  const long long row = a.ld;

  for (long long s = 0; s < a.n_seg; ++s) {
    if (a.seg_t) t = a.seg_t[s];
    if (out && (!a.seg_obs || a.seg_obs[s])) {
#pragma unroll
      for (int v = 0; v < N; ++v) odeintr_b200::store_stream(out + v * row, x[v]);
      out += N * row;
    }
    const long long nfull = a.seg_nfull ? a.seg_nfull[s] : a.nfull_uniform;
    for (long long k = 0; k < nfull; ++k) {
      stepper.do_step(sys, x, t, dt);
      ++step;
      t = (a.time_rule == ODEINTR_TIME_CONST) ? a.t0 + static_cast<double>(step) * dt : t + dt;
    }
    if (a.seg_hrem) {  // remainder step that lands exactly on the next observation / end time
      const double h = a.seg_hrem[s];
      if (h != 0.0) {
        stepper.do_step(sys, x, t, h);
        t += h;
      }
    }
  }
  if (out && a.record_final) {
#pragma unroll
    for (int v = 0; v < N; ++v) odeintr_b200::store_stream(out + v * row, x[v]);
  }
#pragma unroll
  for (int v = 0; v < N; ++v) a.x[v * a.ld + i] = x[v];
}
