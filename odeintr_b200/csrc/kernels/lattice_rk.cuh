// odeintr_b200/csrc/kernels/lattice_rk.cuh
// K4: fused Runge-Kutta stage kernel for large coupled systems (compile_sys_openmp; the reference runs
// Boost's steppers over std::vector<double> with openmp_range_algebra, R/utils.R:34-38).
//
// One launch = one RHS evaluation over the whole lattice PLUS the stepper's vector algebra for that stage,
// applied while the cell's derivative is still in a register:
//     k   = f(xs)[cell]
//     xt  = x   + a_dt * k      (input state of the next stage)          -- skipped on the last stage
//     acc = acc + b_dt * k      (running weighted sum  x + b1 k1 + b2 k2 + ...)
// The running sum is evaluated left to right exactly like Boost's scale_sumN, so results are bit-identical
// to the reference's unfused  k-vectors-then-sum  schedule, with 128 B/cell/step of HBM traffic for rk4
// instead of 208 B (+64 B of by-value vector copies), SURVEY.md §8(d).
//
// Memory-bound: every thread reads 8-24 B and writes 8-16 B per cell, fully coalesced (consecutive threads
// own consecutive cells).  Every block owns one contiguous chunk of cells (page / TLB locality: an SM touches
// 1/gridDim of each array instead of striding over all of it); stencil neighbours are served by L1/L2
// through the read-only path.  Two entry points: the single-GPU kernel indexes the arrays directly, the
// sharded one maps global indices into the rank's slab + ghost window.
#pragma once
#include "odeintr_b200/lattice_view.cuh"

namespace odeintr_b200 {

template <bool SHARDED, bool HAS_XT>
ODEINTR_DEV void lattice_stage_body(const odeintr_lattice_args& a) {
  typedef odeintr::system_type_t<lattice_view_t<SHARDED> > Sys;
  constexpr int F = ODEINTR_NFIELDS;
  Sys sys;
  sys.load_params(a.pars, 1, 0);
  long long off[F];
  sys.field_offsets(off);

  lattice_view_t<SHARDED> xs;
  xs.p = a.xs; xs.n_total = a.n_total; xs.n_local = a.n; xs.c0 = a.c0; xs.ghost = a.ghost; xs.pitch = a.pitch;

  // local position of entry (field f, cell c) is base[f] + c: fields are `pitch` apart, cell c sits at c - c0
  long long base[F];
#pragma unroll
  for (int f = 0; f < F; ++f) base[f] = SHARDED ? (off[f] / a.n_total) * a.pitch - a.c0 : off[f];

  // cells this launch computes: the user loop's range (unsharded) or this rank's window [lo, hi) of global
  // cells, which may reach below 0 / above n_total: those are periodic images and are computed too (they are
  // the ghost cells the next stage reads); split into one contiguous chunk per block (a multiple of the
  // block size, so rows stay aligned)
  const long long first = SHARDED ? a.lo : (sys.cell_start() > a.lo ? sys.cell_start() : a.lo);
  const long long last = SHARDED ? a.hi : (sys.cell_bound() < a.hi ? sys.cell_bound() : a.hi);
  if (last <= first) return;
  // chunks start at multiples of 32 cells counted from local cell 0 (which the host keeps 128-byte aligned),
  // so every warp reads and writes whole 128-byte lines even when the window starts a few ghost cells early
  long long origin = first;
  if (SHARDED) {
    const long long rel = first - a.c0;
    origin = a.c0 + (rel >= 0 ? rel / 32 : -((-rel + 31) / 32)) * 32;
  }
  long long chunk = (last - origin + gridDim.x - 1) / gridDim.x;
  chunk = (chunk + blockDim.x - 1) / blockDim.x * blockDim.x;
  const long long c_lo = origin + (long long)blockIdx.x * chunk;
  long long c_hi = c_lo + chunk;
  if (c_hi > last) c_hi = last;
  for (long long c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
    if (SHARDED && c < first) continue;
    double k[F];
#pragma unroll
    for (int f = 0; f < F; ++f) k[f] = 0.0;
    long long cg = c;
    if (SHARDED) {
      if (cg < 0) cg += a.n_total; else if (cg >= a.n_total) cg -= a.n_total;
      if (cg < sys.cell_start() || cg >= sys.cell_bound()) continue;  // not a cell of the user's loop
    }
    sys.cell(xs, k, cg, a.t);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const long long idx = base[f] + c;
      if (HAS_XT) a.xt[idx] = a.x[idx] + a.a_dt * k[f];
      a.acc_out[idx] = a.acc_in[idx] + a.b_dt * k[f];
    }
  }
}

}  // namespace odeintr_b200

extern "C" __global__ void __launch_bounds__(256)
odeintr_lattice_stage(const odeintr_lattice_args a) {
  if (a.xt) odeintr_b200::lattice_stage_body<false, true>(a);
  else odeintr_b200::lattice_stage_body<false, false>(a);
}

extern "C" __global__ void __launch_bounds__(256)
odeintr_lattice_stage_sharded(const odeintr_lattice_args a) {
  if (a.xt) odeintr_b200::lattice_stage_body<true, true>(a);
  else odeintr_b200::lattice_stage_body<true, false>(a);
}
