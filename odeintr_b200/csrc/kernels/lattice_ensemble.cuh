// odeintr_b200/csrc/kernels/lattice_ensemble.cuh
// K5: ensembles of wider systems -- one WARP (up to 32 cells per field) or one BLOCK (up to 8192 states) per
// trajectory.
//
// One trajectory per thread (K1-K3) stops paying when the state no longer fits a thread's registers.  A wider system
// can only be spread over the lanes of a warp when its components run the same code, i.e. when the snippet is a loop
// over cells -- the compile_sys_openmp form (/root/reference R/odeintr.R:396-405).  An ensemble of such systems
// (a parameter sweep of a discretised PDE, a batch of initial conditions of an oscillator chain: in the reference an
// R loop around NAME_set_params() + NAME(), README.md:157-163) runs here with one trajectory per block:
//   * the block keeps the trajectory on chip for the WHOLE integrate call: every thread owns CPT cells whose state,
//     running sum and stage state live in registers; stage states also go to two shared-memory rings (with `halo`
//     ghost cells holding the periodic images) where the neighbours read them -- the temporally blocked step of
//     lattice_tiled.cuh with the tile being the whole lattice, iterated over all steps;
//   * HBM sees the initial state, the observer rows and the final state, nothing per step;
//   * the step schedule is the host-computed one of K1 (Boost's integrate_const / integrate_times / integrate_adaptive
//     loops for plain steppers), so every trajectory takes exactly the reference's step sequence.
// Fixed-step rk4, bit-identical to running the trajectories one at a time (and to the oracle).
// Threads per trajectory: the cells of one field rounded up to whole warps, at most 256 (512 beyond 2048 cells).
// A lattice of up to 32 cells per field is literally one warp per trajectory: four such warps share a block and
// synchronise with __syncwarp instead of the block barrier.
#pragma once
#include "odeintr_b200/lattice_tiled.cuh"

#define ODEINTR_ENS_CELLS (ODEINTR_SYS_SIZE / ODEINTR_NFIELDS)
#if ODEINTR_ENS_CELLS > 2048
#define ODEINTR_ENS_THREADS 512
#elif ODEINTR_ENS_CELLS >= 256
#define ODEINTR_ENS_THREADS 256
#else
#define ODEINTR_ENS_THREADS ((ODEINTR_ENS_CELLS + 31) / 32 * 32)
#endif
#define ODEINTR_ENS_CPT ((ODEINTR_ENS_CELLS + ODEINTR_ENS_THREADS - 1) / ODEINTR_ENS_THREADS)
#if ODEINTR_ENS_THREADS == 32
#define ODEINTR_ENS_GROUP 4  // trajectories (warps) per block
#else
#define ODEINTR_ENS_GROUP 1
#endif

// registers: 3 doubles per owned state (x, running sum, stage state)
#if !defined(ODEINTR_LATTICE_2D) && !defined(ODEINTR_NO_SYMBOLIC) && ODEINTR_LATTICE_STAGES == 4 && ODEINTR_ENS_CPT * ODEINTR_NFIELDS <= 16 && \
    ODEINTR_ENS_CELLS >= 3

extern "C" __global__ void __launch_bounds__(ODEINTR_ENS_THREADS * ODEINTR_ENS_GROUP)
odeintr_lattice_ensemble(const odeintr_lattice_ensemble_args a) {
  using namespace odeintr_b200;
  typedef lattice_rel_view<true> View;
  typedef odeintr::system_type_t<View> Sys;
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int L = ODEINTR_ENS_CELLS;
  constexpr int NT = ODEINTR_ENS_THREADS;
  constexpr int C = ODEINTR_ENS_CPT;
  constexpr int G = ODEINTR_ENS_GROUP;
  extern __shared__ double smem[];
  const int tid = threadIdx.x % NT, grp = threadIdx.x / NT;
  const int h = (int)a.halo;
  const int W = L + 2 * h;
  double* const A = smem + grp * (2 * F * W);  // x at the start of a step, then the stage-2 state
  double* const B = A + F * W;                 // stage-1 state, then the stage-3 state
  const long long inst = a.first + (long long)blockIdx.x * G + grp;
  const long long ld = a.p.ld;
  if (inst >= a.p.m) return;  // only whole warps of a grouped block (they never meet a block barrier)
  // the threads of one trajectory wait for each other: a warp barrier when the trajectory is one warp
  auto sync = [&]() {
    if (NT == 32) __syncwarp();
    else __syncthreads();
  };

  Sys sys;
  sys.load_params(a.p.pars, a.p.par_ld, inst * a.p.par_inc);
  long long off[F];
  sys.field_offsets(off, 0);
  int fld[F];
#pragma unroll
  for (int f = 0; f < F; ++f) fld[f] = static_cast<int>(off[f] / L);
  const long long start = sys.cell_start(), bound = sys.cell_bound();

  // entry (field q, cell j) of a ring; cells 0 .. h-1 and L-h .. L-1 are mirrored into the ghost cells
  auto put = [&](double* ring, const int q, const int j, const double v) {
    double* r = ring + q * W + h;
    r[j] = v;
    if (j < h) r[L + j] = v;
    if (j >= L - h) r[j - L] = v;
  };

  double xs[C][F], acc[C][F], own[C][F];
  bool live[C];  // a cell of the user's loop (the others keep their value)
#pragma unroll
  for (int k = 0; k < C; ++k) {
    const int j = tid + k * NT;
    live[k] = j < L && j >= start && j < bound;
#pragma unroll
    for (int f = 0; f < F; ++f) {
      double v = 0.0;
      if (j < L) {
        v = a.p.x[(static_cast<long long>(fld[f]) * L + j) * ld + inst];
        put(A, fld[f], j, v);
        put(B, fld[f], j, v);
      }
      xs[k][f] = v;
      acc[k][f] = v;
    }
  }
  sync();

  View view;
  view.n_total = L; view.width = W; view.halo = h; view.bad = 0;
  rel_index idx;
  idx.off = 0; idx.cells = L; idx.rel = true;

  auto do_step = [&](const double t, const double dt) {
    const double half = 1.0 / 2.0, b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
#pragma unroll
    for (int k = 0; k < C; ++k)
#pragma unroll
      for (int f = 0; f < F; ++f) own[k][fld[f]] = xs[k][f];
#pragma unroll
    for (int stage = 1; stage <= 4; ++stage) {
      const double* src = (stage & 1) ? A : B;
      double* dst = (stage & 1) ? B : A;  // stage 4 leaves the new state in A, the next step's input
      const double a_dt = stage == 1 ? half * dt : (stage == 2 ? half * dt : 1.0 * dt);
      const double b_dt = stage == 1 ? b1 * dt : (stage == 2 ? b2 * dt : (stage == 3 ? b2 * dt : b1 * dt));
      const double ts = stage == 1 ? t : (stage == 4 ? t + 1.0 * dt : t + half * dt);
#pragma unroll
      for (int k = 0; k < C; ++k) {
        if (!live[k]) continue;
        const int j = tid + k * NT;
        double kk[F];
#pragma unroll
        for (int f = 0; f < F; ++f) kk[f] = 0.0;
        view.p = src + h + j;
        view.own = own[k];
        view.centre = idx.centre = j;
        sys.cell(view, kk, idx, ts);
#pragma unroll
        for (int f = 0; f < F; ++f) {
          acc[k][f] = (stage == 1 ? xs[k][f] : acc[k][f]) + b_dt * kk[f];  // the running sum starts from x
          const double v = stage < 4 ? xs[k][f] + a_dt * kk[f] : acc[k][f];
          put(dst, fld[f], j, v);
          own[k][fld[f]] = v;
          if (stage == 4) xs[k][f] = v;
        }
      }
      sync();
    }
  };

  // the schedule of odeintr_plain_ensemble (ensemble_plain.cuh): observe, full steps, remainder step, per segment
  const double dt = a.p.dt;
  double t = a.p.t0;
  long long step = 0;
  double* out = a.p.out ? a.p.out + inst : nullptr;
  auto observe = [&]() {
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const int j = tid + k * NT;
      if (j < L) {
#pragma unroll
        for (int f = 0; f < F; ++f) store_stream(out + (static_cast<long long>(fld[f]) * L + j) * ld, xs[k][f]);
      }
    }
    out += static_cast<long long>(ODEINTR_SYS_SIZE) * ld;
  };
  for (long long s = 0; s < a.p.n_seg; ++s) {
    if (a.p.seg_t) t = a.p.seg_t[s];
    if (out && (!a.p.seg_obs || a.p.seg_obs[s])) observe();
    const long long nfull = a.p.seg_nfull ? a.p.seg_nfull[s] : a.p.nfull_uniform;
    for (long long q = 0; q < nfull; ++q) {
      do_step(t, dt);
      ++step;
      t = (a.p.time_rule == ODEINTR_TIME_CONST) ? a.p.t0 + static_cast<double>(step) * dt : t + dt;
    }
    if (a.p.seg_hrem) {  // remainder step that lands exactly on the next observation / end time
      const double hr = a.p.seg_hrem[s];
      if (hr != 0.0) {
        do_step(t, hr);
        t += hr;
      }
    }
  }
  if (out && a.p.record_final) observe();
#pragma unroll
  for (int k = 0; k < C; ++k) {
    const int j = tid + k * NT;
    if (j < L) {
#pragma unroll
      for (int f = 0; f < F; ++f) a.p.x[(static_cast<long long>(fld[f]) * L + j) * ld + inst] = xs[k][f];
    }
  }
  if (view.bad) *a.flag = 1;
}

#endif
