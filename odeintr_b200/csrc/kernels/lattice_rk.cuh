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
// own consecutive cells); stencil neighbours are served by L1/L2 through the read-only path.
#pragma once
#include "odeintr_b200/lattice_view.cuh"

extern "C" __global__ void __launch_bounds__(256)
odeintr_lattice_stage(const odeintr_lattice_args a) {
  typedef odeintr::system_type Sys;
  constexpr int F = ODEINTR_NFIELDS;
  Sys sys;
  sys.load_params(a.pars, 1, 0);
  long long off[F];
  sys.field_offsets(off);

  odeintr_b200::lattice_view xs;
  xs.p = a.xs; xs.n_total = a.n_total; xs.n_local = a.n; xs.c0 = a.c0; xs.ghost = a.ghost; xs.pitch = a.pitch;
  xs.sharded = a.sharded != 0;

  // cells this launch computes: the user loop's range, clipped to this rank's window [lo, hi) (global cells)
  // local position of entry (field f, cell c) is base[f] + c: fields are `pitch` apart, cell c sits at c - c0
  long long base[F];
#pragma unroll
  for (int f = 0; f < F; ++f) base[f] = a.sharded ? (off[f] / a.n_total) * a.pitch - a.c0 : off[f];

  const long long first = sys.cell_start() > a.lo ? sys.cell_start() : a.lo;
  const long long last = sys.cell_bound() < a.hi ? sys.cell_bound() : a.hi;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long c = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; c < last; c += stride) {
    double k[F];
#pragma unroll
    for (int f = 0; f < F; ++f) k[f] = 0.0;
    // periodic images: on a sharded run the window may start below 0 or end above n_total
    long long cg = c;
    if (cg < 0) cg += a.n_total; else if (cg >= a.n_total) cg -= a.n_total;
    sys.cell(xs, k, cg, a.t);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const long long idx = base[f] + c;
      if (a.xt) a.xt[idx] = a.x[idx] + a.a_dt * k[f];
      a.acc_out[idx] = a.acc_in[idx] + a.b_dt * k[f];
    }
  }
}
