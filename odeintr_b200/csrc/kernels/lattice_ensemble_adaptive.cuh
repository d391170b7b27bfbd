// odeintr_b200/csrc/kernels/lattice_ensemble_adaptive.cuh
// K5a: ensembles of WIDER systems under the adaptive steppers -- compile_sys_openmp's default method rk5_i
// (dense-output dopri5), rk5_a, rk54_a, rk78_a, bs and bsd --, one WARP or one BLOCK per trajectory (lattice_ensemble.cuh has
// the fixed-step rk4 form).  In the reference: an R loop around NAME_set_params() + NAME() / NAME_at()
// (README.md:157-163) over a system compiled with compile_sys_openmp's defaults (R/odeintr.R:418-421).
//
// The threads of a trajectory run Boost's integrate_const / integrate_times / integrate_adaptive loops
// (integrate_loops.cuh) and the per-trajectory steppers of adaptive_steppers.cuh UNCHANGED -- the same code every
// thread of K2 runs for its own small system -- on a distributed state: a thread's "state vector" holds the entries of
// its own cells (CPT cells x fields, in registers, k1 .. k7 and the dense-output history included).  Two things make
// the distributed vectors behave like the whole one:
//   * the right-hand side publishes the stage state to a shared-memory ring (ghost cells hold the periodic images),
//     waits for the trajectory's threads and evaluates the user's loop body for the thread's cells (two rings
//     alternate, so one barrier per evaluation is enough);
//   * the error norm is the maximum over the trajectory's threads (shuffles, and shared memory between warps), handed
//     back to every thread: all of them take the controller's decisions with the same numbers, so step sizes, accepted /
//     rejected counts and the control flow are uniform across the trajectory.
// HBM sees the initial state, the requested samples and the final state.  Every trajectory follows the oracle's step
// sequence up to the controller's pow() (libdevice vs glibc), like K2: results within 10 x rtol, in practice 1e-13.
#pragma once
#include "odeintr_b200/lattice_ensemble.cuh"

// registers: ~13 doubles per owned entry (x, k1, x_new, k7, x_old, k1_old, k2 .. k6, stage state, error estimate)
#if !defined(ODEINTR_LATTICE_2D) && !defined(ODEINTR_NO_SYMBOLIC) && \
    (ODEINTR_LATTICE_STAGES == 7 || ODEINTR_LATTICE_STAGES == 6 || ODEINTR_LATTICE_STAGES == 13 || ODEINTR_LATTICE_STAGES == 2 || \
     ODEINTR_LATTICE_STAGES == 1) && \
    ODEINTR_ENS_CPT * ODEINTR_NFIELDS <= 4 && ODEINTR_ENS_CELLS >= 3 && ODEINTR_ENS_CELLS <= 2048
#define ODEINTR_HAVE_ENS_ADAPTIVE 1
#include "odeintr_b200/integrate_device.cuh"
#include "odeintr_b200/steppers_plain.cuh"

namespace odeintr_b200 {

// the user's system as the steppers see it: vec<NE> holds the calling thread's entries of a distributed vector
template <class UserSys>
struct block_system {
  static constexpr int F = ODEINTR_NFIELDS;
  static constexpr int L = ODEINTR_ENS_CELLS;
  static constexpr int NT = ODEINTR_ENS_THREADS;
  static constexpr int C = ODEINTR_ENS_CPT;
  static constexpr int NE = C * F;
  typedef lattice_rel_view<true> View;
  UserSys user;
  int bad;           // a read beyond the stencil radius was seen
  double* ring[2];   // stage-state rings of this trajectory (fields x (L + 2 halo))
  double* red;       // per-warp partial maxima, two alternating rows
  int phase, rphase;
  int tid, h, W;
  int fld[F];
  bool live[C];

  ODEINTR_DEV void barrier() const {
    if (NT == 32) __syncwarp();
    else __syncthreads();
  }
  // The steppers ask for the right-hand side through vec<NE> references; it takes and returns the thread's entries by
  // value so that the stage vectors stay in registers.  (History, profiles/r02_k5a.md: first behind a call because the
  // inlined form compiled for minutes -- until ncu showed 12 800 instructions per warp and attempt, 3 % of them FP64: the
  // view's lattice length was a run-time member, so every state access paid 64-bit divisions.  With the view rebuilt
  // from constants below the inlined form compiles in seconds and runs 7x faster than that first version.)
  ODEINTR_DEV void sys(const vec<NE>& x, vec<NE>& k, const double t) { k = rhs(x, t); }
  ODEINTR_DEV vec<NE> rhs(const vec<NE> x, const double t) {
    vec<NE> k;
    double* const R = ring[phase];
    phase ^= 1;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = tid + c * NT;
      if (j < L) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
          double* r = R + fld[f] * W + h;
          const double v = x[c * F + f];
          r[j] = v;
          if (j < h) r[L + j] = v;
          if (j >= L - h) r[j - L] = v;
        }
      }
    }
    barrier();
    // the view and the symbolic index are built here from compile-time constants: the snippet's index expressions then
    // fold to fixed shared-memory offsets (a run-time lattice length costs 64-bit divisions per state access)
    View view;
    view.n_total = L; view.width = W; view.halo = h; view.bad = 0;
    rel_index idx;
    idx.off = 0; idx.cells = L; idx.rel = true;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = tid + c * NT;
      double kk[F], own[F];
#pragma unroll
      for (int f = 0; f < F; ++f) { kk[f] = 0.0; own[fld[f]] = x[c * F + f]; }
      if (live[c]) {
        view.p = R + h + j;
        view.own = own;
        view.centre = idx.centre = j;
        user.cell(view, kk, idx, t);
      }
#pragma unroll
      for (int f = 0; f < F; ++f) k[c * F + f] = kk[f];
    }
    bad |= view.bad;
    return k;
  }
  // maximum over the trajectory's threads, known to all of them afterwards
  ODEINTR_DEV double block_max(double e) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, s));
    if (NT == 32) return e;
    double* const row = red + rphase * (NT / 32);
    rphase ^= 1;
    if ((tid & 31) == 0) row[tid >> 5] = e;
    __syncthreads();
    double m = row[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) m = fmax(m, row[w]);
    return m;  // (the other row is written next: no second barrier)
  }
};

// observer of a distributed state: every thread stores its own entries of the sample row
template <int NE>
struct block_sample_writer {
  double* out;   // this instance's column of sample 0, or null
  double* t_rec; // this instance's sample times (ragged records), or null
  long long ld;
  int rows, max_rows;
  int tid;
  const int* fld;
  ODEINTR_DEV void operator()(const vec<NE>& x, const double t) {
    constexpr int F = ODEINTR_NFIELDS, L = ODEINTR_ENS_CELLS, NT = ODEINTR_ENS_THREADS, C = ODEINTR_ENS_CPT;
    if (out && rows < max_rows) {
      double* row = out + static_cast<long long>(rows) * ODEINTR_SYS_SIZE * ld;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int j = tid + c * NT;
        if (j < L) {
#pragma unroll
          for (int f = 0; f < F; ++f) store_stream(row + (static_cast<long long>(fld[f]) * L + j) * ld, x[c * F + f]);
        }
      }
      if (t_rec && tid == 0) store_stream(t_rec + static_cast<long long>(rows) * ld, t);
    }
    ++rows;
  }
};

template <template <class, int> class Stepper>
ODEINTR_DEV void lattice_ensemble_adaptive_body(const odeintr_lattice_ensemble_adaptive_args& b) {
  typedef odeintr::system_type_t<lattice_rel_view<true> > UserSys;
  typedef block_system<UserSys> Sys;
  constexpr int F = Sys::F, L = Sys::L, NT = Sys::NT, C = Sys::C, NE = Sys::NE;
  constexpr int G = ODEINTR_ENS_GROUP;
  typedef Stepper<Sys, NE> St;
  const odeintr_adaptive_args& a = b.p;
  extern __shared__ double smem[];
  const int tid = threadIdx.x % NT, grp = threadIdx.x / NT;
  const int h = static_cast<int>(b.halo);
  const int W = L + 2 * h;
  const long long inst = static_cast<long long>(blockIdx.x) * G + grp;
  if (inst >= a.m) return;  // only whole warps of a grouped block (they never meet a block barrier)

  Sys sys;
  sys.user.load_params(a.pars, a.par_ld, inst * a.par_inc);
  sys.ring[0] = smem + grp * (2 * F * W + 2 * (NT / 32));
  sys.ring[1] = sys.ring[0] + F * W;
  sys.red = sys.ring[1] + F * W;
  sys.phase = 0; sys.rphase = 0;
  sys.tid = tid; sys.h = h; sys.W = W;
  {
    long long off[F];
    sys.user.field_offsets(off, 0);
#pragma unroll
    for (int f = 0; f < F; ++f) sys.fld[f] = static_cast<int>(off[f] / L);
  }
  const long long start = sys.user.cell_start(), bound = sys.user.cell_bound();
  sys.bad = 0;

  vec<NE> x;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = tid + c * NT;
    sys.live[c] = j < L && j >= start && j < bound;
#pragma unroll
    for (int f = 0; f < F; ++f)
      x[c * F + f] = j < L ? a.x[(static_cast<long long>(sys.fld[f]) * L + j) * a.ld + inst] : 0.0;
  }

  St st;
  st.setup(a.atol, a.rtol);
  st.counters().budget = a.max_tries;
  block_sample_writer<NE> obs;
  obs.out = a.out ? a.out + inst : nullptr;
  obs.t_rec = a.t_rec ? a.t_rec + inst : nullptr;
  obs.ld = a.ld; obs.rows = 0; obs.max_rows = a.max_rows; obs.tid = tid; obs.fld = sys.fld;

  bool ok;
  if constexpr (St::is_dense) {
    if (a.mode == ODEINTR_MODE_TIMES) ok = integrate_times_dense(st, sys, x, a.times, a.n_times, a.dt, obs);
    else if (a.mode == ODEINTR_MODE_CONST) ok = integrate_const_dense(st, sys, x, a.t0, a.t1, a.dt, obs);
    else ok = integrate_adaptive_dense(st, sys, x, a.t0, a.t1, a.dt, obs);
  } else {
    double dt = a.dt;
    if (a.mode == ODEINTR_MODE_TIMES) ok = integrate_times_controlled(st, sys, x, a.times, a.n_times, dt, obs);
    else if (a.mode == ODEINTR_MODE_CONST) ok = integrate_const_controlled(st, sys, x, a.t0, a.t1, dt, obs);
    else {
      double t = a.t0;
      ok = integrate_adaptive_controlled(st, sys, x, t, a.t1, dt, obs);
    }
  }
  (void)ok;  // the status word carries the failure

  bool finite = true;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = tid + c * NT;
    if (j < L) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        a.x[(static_cast<long long>(sys.fld[f]) * L + j) * a.ld + inst] = x[c * F + f];
        finite = finite && isfinite(x[c * F + f]);
      }
    }
  }
  step_counters& cnt = st.counters();
  const double worst = sys.block_max(finite ? 0.0 : 1.0);
  if (tid == 0) {
    if (a.accepted) a.accepted[inst] = cnt.accepted;
    if (a.rejected) a.rejected[inst] = cnt.rejected;
    int status = cnt.status;
    if (worst > 0.0 && status == ODEINTR_STATUS_OK) status = ODEINTR_STATUS_NONFINITE;
    if (a.status) a.status[inst] = status;
    if (status != ODEINTR_STATUS_OK && a.fail_flags) atomicOr(a.fail_flags, status);
    if (a.rows) a.rows[inst] = obs.rows;
    if (a.rows_min) atomicMin(a.rows_min, obs.rows);
  }
  if (sys.bad) *b.flag = 1;
}

// The fixed-step methods other than rk4 (euler, rk54, rk5, rk78; rk4 has the register-blocked kernel of
// lattice_ensemble.cuh) on the same distributed state: K1's host-computed step schedule, the plain steppers of K1.
template <template <class, int> class Stepper>
ODEINTR_DEV void lattice_ensemble_plain_body(const odeintr_lattice_ensemble_args& b) {
  typedef odeintr::system_type_t<lattice_rel_view<true> > UserSys;
  typedef block_system<UserSys> Sys;
  constexpr int F = Sys::F, L = Sys::L, NT = Sys::NT, C = Sys::C, NE = Sys::NE;
  constexpr int G = ODEINTR_ENS_GROUP;
  const odeintr_plain_args& a = b.p;
  extern __shared__ double smem[];
  const int tid = threadIdx.x % NT, grp = threadIdx.x / NT;
  const int h = static_cast<int>(b.halo);
  const int W = L + 2 * h;
  const long long inst = b.first + static_cast<long long>(blockIdx.x) * G + grp;
  if (inst >= a.m) return;

  Sys sys;
  sys.user.load_params(a.pars, a.par_ld, inst * a.par_inc);
  sys.ring[0] = smem + grp * (2 * F * W + 2 * (NT / 32));
  sys.ring[1] = sys.ring[0] + F * W;
  sys.red = sys.ring[1] + F * W;
  sys.phase = 0; sys.rphase = 0;
  sys.tid = tid; sys.h = h; sys.W = W;
  {
    long long off[F];
    sys.user.field_offsets(off, 0);
#pragma unroll
    for (int f = 0; f < F; ++f) sys.fld[f] = static_cast<int>(off[f] / L);
  }
  const long long start = sys.user.cell_start(), bound = sys.user.cell_bound();
  sys.bad = 0;
  vec<NE> x;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = tid + c * NT;
    sys.live[c] = j < L && j >= start && j < bound;
#pragma unroll
    for (int f = 0; f < F; ++f)
      x[c * F + f] = j < L ? a.x[(static_cast<long long>(sys.fld[f]) * L + j) * a.ld + inst] : 0.0;
  }
  Stepper<Sys, NE> stepper;
  stepper.reset();
  block_sample_writer<NE> obs;
  obs.out = a.out ? a.out + inst : nullptr;
  obs.t_rec = nullptr;
  obs.ld = a.ld; obs.rows = 0; obs.max_rows = 0x7fffffff; obs.tid = tid; obs.fld = sys.fld;

  // the schedule of odeintr_plain_ensemble (ensemble_plain.cuh): observe, full steps, remainder step, per segment
  const double dt = a.dt;
  double t = a.t0;
  long long step = 0;
  for (long long s = 0; s < a.n_seg; ++s) {
    if (a.seg_t) t = a.seg_t[s];
    if (obs.out && (!a.seg_obs || a.seg_obs[s])) obs(x, t);
    const long long nfull = a.seg_nfull ? a.seg_nfull[s] : a.nfull_uniform;
    for (long long q = 0; q < nfull; ++q) {
      stepper.do_step(sys, x, t, dt);
      ++step;
      t = (a.time_rule == ODEINTR_TIME_CONST) ? a.t0 + static_cast<double>(step) * dt : t + dt;
    }
    if (a.seg_hrem) {  // remainder step that lands exactly on the next observation / end time
      const double hr = a.seg_hrem[s];
      if (hr != 0.0) {
        stepper.do_step(sys, x, t, hr);
        t += hr;
      }
    }
  }
  if (obs.out && a.record_final) obs(x, t);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = tid + c * NT;
    if (j < L) {
#pragma unroll
      for (int f = 0; f < F; ++f) a.x[(static_cast<long long>(sys.fld[f]) * L + j) * a.ld + inst] = x[c * F + f];
    }
  }
  if (sys.bad) *b.flag = 1;
}

}  // namespace odeintr_b200

#if ODEINTR_LATTICE_STAGES != 2  // (bs / bsd are never plain steppers)
extern "C" __global__ void __launch_bounds__(ODEINTR_ENS_THREADS * ODEINTR_ENS_GROUP)
odeintr_lattice_ensemble_plain(const odeintr_lattice_ensemble_args a) {
#if ODEINTR_LATTICE_STAGES == 1
  odeintr_b200::lattice_ensemble_plain_body<odeintr_b200::euler_stepper>(a);
#elif ODEINTR_LATTICE_STAGES == 6
  odeintr_b200::lattice_ensemble_plain_body<odeintr_b200::rk54_stepper>(a);
#elif ODEINTR_LATTICE_STAGES == 7
  odeintr_b200::lattice_ensemble_plain_body<odeintr_b200::rk5_stepper>(a);
#else
  odeintr_b200::lattice_ensemble_plain_body<odeintr_b200::rk78_stepper>(a);
#endif
}
#endif

// one kernel per stepper the translation unit's method can ask for (the stage count is how the generator names it)
#ifndef ODEINTR_ENS_ADAPTIVE_MIN_BLOCKS  // resident blocks per SM the adaptive kernels are compiled for (register cap)
#define ODEINTR_ENS_ADAPTIVE_MIN_BLOCKS 2  // measured (profiles/r02_k5a.md): 128 registers / 2 blocks 1.89 ms, 156 / 1 block 2.73 ms
#endif
#define ODEINTR_ENS_ADAPTIVE_KERNEL(NAME, STEPPER)                                                  \
  extern "C" __global__ void __launch_bounds__(ODEINTR_ENS_THREADS * ODEINTR_ENS_GROUP, ODEINTR_ENS_ADAPTIVE_MIN_BLOCKS) \
  NAME(const odeintr_lattice_ensemble_adaptive_args a) {                                            \
    odeintr_b200::lattice_ensemble_adaptive_body<odeintr_b200::STEPPER>(a);                         \
  }
#if ODEINTR_LATTICE_STAGES == 7    // rk5_i, rk5_a
ODEINTR_ENS_ADAPTIVE_KERNEL(odeintr_lattice_ensemble_dopri5_dense, dopri5_dense)
ODEINTR_ENS_ADAPTIVE_KERNEL(odeintr_lattice_ensemble_dopri5_controlled, dopri5_controlled)
#elif ODEINTR_LATTICE_STAGES == 6  // rk54_a
ODEINTR_ENS_ADAPTIVE_KERNEL(odeintr_lattice_ensemble_rk54_controlled, rk54_controlled)
#elif ODEINTR_LATTICE_STAGES == 13  // rk78_a
ODEINTR_ENS_ADAPTIVE_KERNEL(odeintr_lattice_ensemble_rk78_controlled, rk78_controlled)
#elif ODEINTR_LATTICE_STAGES == 2  // bs, bsd
ODEINTR_ENS_ADAPTIVE_KERNEL(odeintr_lattice_ensemble_bs_controlled, bs_controlled)
ODEINTR_ENS_ADAPTIVE_KERNEL(odeintr_lattice_ensemble_bsd_dense, bsd_dense)
#endif
#undef ODEINTR_ENS_ADAPTIVE_KERNEL

#endif  // adaptive ensembles of wider systems
