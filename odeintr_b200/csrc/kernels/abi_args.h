// odeintr_b200/csrc/kernels/abi_args.h
// Plain-old-data launch descriptors shared by the host C-ABI (abi.cu) and the NVRTC-compiled
// stepper kernels.  Only builtin types: this header is compiled by nvcc (host side) and by
// NVRTC (device side, no host headers available).
#pragma once

// How the stepper's time variable advances between steps.
//   ODEINTR_TIME_CONST : t = t0 + step*dt   (integrate_const, integrate_n_steps)
//   ODEINTR_TIME_ACCUM : t += h, re-based to the requested time at each observation
//                        (integrate_times for plain steppers)
#define ODEINTR_TIME_CONST 0
#define ODEINTR_TIME_ACCUM 1

// Per-instance status words written by the kernels (failure detection).
#define ODEINTR_STATUS_OK 0
#define ODEINTR_STATUS_STEP_OVERFLOW 1   // >500 consecutive rejected tries (Boost step_adjustment_error)
#define ODEINTR_STATUS_NONFINITE 2       // state became NaN/Inf
#define ODEINTR_STATUS_STEP_BUDGET 4     // more than max_tries step attempts in one call (guards the GPU against
                                         // the endless loops Boost itself can enter, e.g. dt underflow at a blow-up)

// Fixed-step ("plain stepper") ensemble launch: one trajectory per thread.
// State and samples are structure-of-arrays with the instance index fastest:
//   x   [v * ld + i]                       v < N, i < m
//   out [(s * N + v) * ld + i]             s < n_seg + 1 (sample-major SoA; one coalesced row per variable)
// The step schedule is instance-independent and computed on the host with the reference's own
// loops (less_eq_with_sign etc.), so every thread executes exactly the reference's step sequence:
//   for s in [0, n_seg):  t = seg_t[s]; observe if seg_obs[s]; nfull(s) steps of dt;
//                         optional remainder step hrem(s)
//   observe (final) if record_final
struct odeintr_plain_args {
  double* x;                  // in/out state, N x ld
  const double* pars;         // P x par_ld (par_inc = 1) or P values broadcast (par_inc = 0)
  double* out;                // samples or null (no_record)
  long long m;                // instances handled by this launch
  long long ld;               // leading dimension of x and out rows (>= m)
  long long par_ld;
  long long par_inc;
  double t0;
  double dt;
  long long n_seg;            // schedule segments
  long long nfull_uniform;    // steps per segment when seg_nfull is null
  const long long* seg_nfull; // per-segment full-step counts or null
  const double* seg_hrem;     // per-segment remainder step (0 = none) or null
  const double* seg_t;        // time at the start of each segment, exactly as the reference's loop
                              // computes it (null: t follows time_rule from t0)
  const unsigned char* seg_obs; // observe at the start of segment s? (null: every segment)
  int time_rule;
  int record_final;           // store the sample after the last segment
};

// Adaptive ensemble launch (controlled / dense-output steppers: dopri5, cash-karp, rosenbrock4).
// The stepper family (controlled vs dense) is fixed at compile time by ODEINTR_STEPPER; `mode` picks the
// integrate_* loop.  A null `out` is Boost's null_observer.
#define ODEINTR_MODE_CONST 0     // integrate_const:    samples at t0 + k*dt up to t1
#define ODEINTR_MODE_TIMES 1     // integrate_times:    samples at times[0..n_times)
#define ODEINTR_MODE_ADAPTIVE 2  // integrate_adaptive: one sample per accepted step (ragged) + the end point

struct odeintr_adaptive_args {
  double* x;                  // in/out state, N x ld
  const double* pars;
  double* out;                // (s * N + v) * ld + i, or null
  double* t_rec;              // per-instance sample times s * ld + i (ragged records) or null
  long long m;
  long long ld;
  long long par_ld;
  long long par_inc;
  double t0;
  double t1;
  double dt;                  // initial step (adaptive) / observation step (const)
  const double* times;        // observation times (times mode)
  long long n_times;
  double atol;
  double rtol;
  int mode;
  int max_rows;               // capacity of out in rows
  long long max_tries;        // per-instance budget of step attempts (accepted + rejected) per call
  long long* accepted;        // per-instance counters (may be null)
  long long* rejected;
  int* status;
  int* rows;                  // per-instance observer calls (may be null)
  int* fail_flags;            // one word: OR of every instance's non-zero status (may be null)
  int* rows_min;              // one word: minimum over instances of the observer calls made (may be null)
  const int* perm;            // K6 instance order: thread j integrates instance perm[j] (null: instance j).  Results
                              // land where the instance lives, so the order changes which trajectories share a warp
                              // and nothing else
};

// Ensemble of wider systems under an adaptive stepper, one warp / block per trajectory (lattice_ensemble_adaptive.cuh)
struct odeintr_lattice_ensemble_adaptive_args {
  odeintr_adaptive_args p;
  long long halo;             // stencil radius (ghost cells per side of the shared-memory rings)
  int* flag;                  // set to 1 when the snippet reaches beyond the halo
};

// Large coupled system ("lattice") fused Runge-Kutta stage launch, SURVEY.md §8(d) config 5:
//   k = f(xs)[cell];  xt = x + a_dt*k (if xt != null);  acc_out = acc_in + b_dt*k
// All pointers address local cell 0 of field 0; on a sharded run ghost cells sit at negative offsets and
// the slabs of consecutive fields are `pitch` doubles apart.
struct odeintr_lattice_args {
  const double* xs;           // stage input state (argument of f)
  const double* x;            // state at the start of the step (base of xt)
  double* xt;                 // next stage's input state, or null on the last stage
  const double* acc_in;       // running sum so far (x itself on the first stage)
  double* acc_out;            // running sum (the new state on the last stage)
  const double* pars;
  long long n;                // cells per field owned by this rank
  long long n_total;          // global cells per field
  long long c0;               // global cell index of local cell 0
  long long ghost;            // ghost cells on each side
  long long pitch;            // doubles between consecutive field slabs (sharded runs)
  long long lo;               // first global cell to compute (may be < 0: periodic image)
  long long hi;               // one past the last global cell to compute
  long long halo;             // stencil radius of the user's loop (sharded runs)
  double t;
  double a_dt;
  double b_dt;
};

// Generic fused stage for large systems under adaptive steppers (lattice_combo.cuh), one GPU:
//   k = f(xs)[cell] (if xs);  v = base + sum c[j]*in[j] (+ ck*k);  out = v;  k_out = k;
//   err_max = max |v| / (eps_abs + eps_rel*(|err_x| + adt*|err_d|))  (if err_max)
struct odeintr_lattice_combo_args {
  const double* xs;
  const double* base;
  const double* in[12];
  double c[12];
  int nin;
  double ck;
  double* out;
  double* k_out;
  const double* err_x;
  const double* err_d;
  double eps_abs, eps_rel, adt;
  unsigned long long* err_max;
  const double* pars;
  long long n_total;
  double t;
  int all;  // also form the sum for the entries no user loop writes (their k is zero): Bulirsch-Stoer's extrapolation
            // tables do not reproduce such entries by x + c * 0
};

// One attempted dopri5 step of a large 1-D system in one launch (lattice_dopri5_tiled.cuh).
struct odeintr_lattice_dopri5_args {
  const double* x_in;         // state at the start of the step
  const double* d_in;         // its derivative k1 = f(x_in, t)
  double* x_out;              // 5th-order solution (must not alias x_in)
  double* d_out;              // k7 = f(x_out, t + dt)
  double* dense_out;          // null, or: evaluate the dense output of this step instead (x_out, d_out and err_max
  double dcoef[6];            //   stay untouched): dense_out = x_in + dcoef . (k1, k3, k4, k5, k6, k7), left to right
  const double* pars;
  long long halo;             // stencil radius
  double t, dt;
  double eps_abs, eps_rel, adt;
  unsigned long long* err_max;  // bit pattern of the max-norm error (atomicMax)
  int* flag;                  // set to 1 when the snippet reaches beyond the halo
  // Slab of a sharded run (one process per GPU): the launch produces the cells [own_lo, own_hi) (global numbering);
  // entry (field f, global cell c) of every array sits at f * pitch + (c - c0), ghost cells -- the ring neighbours'
  // cells, periodic images at the ends of the ring -- at the offsets beyond the owned range.  One GPU: own = [0, L),
  // c0 = 0, pitch = L, sharded = 0 (the window is then loaded with the cell index wrapped periodically).
  long long own_lo, own_hi, c0, pitch;
  int sharded;
  int tma;                    // warp kernel (lattice_dopri5_warp.cuh): 16-byte aligned windows, move them with the TMA engine
  int seg_rows;               // streaming M x M kernel (lattice_dopri5_stream2d.cuh): rows per warp task
  double cf[32];              //   and its coefficient * dt products, formed once on the host (odeintr_dopri5_cf below): as
                              //   kernel parameters they are constant-bank operands, not registers
};
// index of the products in odeintr_lattice_dopri5_args::cf
enum odeintr_dopri5_cf {
  ODEINTR_CF_B21 = 0, ODEINTR_CF_B31, ODEINTR_CF_B32, ODEINTR_CF_B41, ODEINTR_CF_B42, ODEINTR_CF_B43, ODEINTR_CF_B51,
  ODEINTR_CF_B52, ODEINTR_CF_B53, ODEINTR_CF_B54, ODEINTR_CF_B61, ODEINTR_CF_B62, ODEINTR_CF_B63, ODEINTR_CF_B64,
  ODEINTR_CF_B65, ODEINTR_CF_C1, ODEINTR_CF_C3, ODEINTR_CF_C4, ODEINTR_CF_C5, ODEINTR_CF_C6,
  ODEINTR_CF_E1, ODEINTR_CF_E3, ODEINTR_CF_E4, ODEINTR_CF_E5, ODEINTR_CF_E6, ODEINTR_CF_E7,  // error weights * dt, or the dense-output weights
  ODEINTR_CF_T2, ODEINTR_CF_T3, ODEINTR_CF_T4, ODEINTR_CF_T5, ODEINTR_CF_T6                   // stage times
};

// Ensemble of wider systems, one trajectory per block (lattice_ensemble.cuh): the fixed-step schedule of
// odeintr_plain_args, block b integrating instance first + b.
struct odeintr_lattice_ensemble_args {
  odeintr_plain_args p;
  long long first;
  long long halo;             // stencil radius (ghost cells per side of the shared-memory rings)
  int* flag;                  // set to 1 when the snippet reaches beyond the halo
};

// One rk4 step of a large 1-D system in one launch, temporally blocked in shared memory (lattice_tiled.cuh).
struct odeintr_lattice_tiled_args {
  const double* x_in;         // state at the start of the step (local cell 0 of field 0)
  double* x_out;              // state after the step (must not alias x_in)
  const double* pars;
  long long n;                // cells per field owned by this rank
  long long n_total;          // global cells per field
  long long c0;               // global cell index of local cell 0
  long long ghost;            // ghost cells on each side of the local slab (sharded)
  long long pitch;            // doubles between field slabs (sharded)
  long long lo;               // owned cells [lo, hi) in global cell numbering
  long long hi;
  long long halo;             // stencil radius declared by the user
  double t;
  double dt;
  int sharded;
  int* flag;                  // set to 1 when the snippet reaches beyond the declared halo
  int tile_first;             // K4t interior kernels: block b computes the tile [lo + (tile_first + b) * T, ... + T),
  int tile_skip;              // tile_skip tiles in all (the persistent TMA kernel walks over them)
  long long lo2;              // general kernel: a second range of cells [lo2, hi2) after [lo, hi) -- what is left on
  long long hi2;              // either side of the cells the interior / warp kernel produces
  int tma;                    // warp kernel: windows are 16-byte aligned, move them with the TMA engine
  int nsub;                   // warp kernel: rk4 steps per pass (1, or 2: the second one at time t2)
  double t2;
};
