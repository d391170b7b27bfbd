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

// cells [c_lo, c_hi) of one block through one view type; entry (f, c) of every array sits at base[f] + c
template <int MODE, bool HAS_XT, int F>
ODEINTR_DEV void lattice_chunk(const odeintr_lattice_args& a, const lattice_view_t<MODE>& xs, const long long* base,
                               const long long c_lo, const long long c_hi, const long long first) {
  typedef odeintr::system_type_t<lattice_view_t<MODE> > Sys;
  constexpr bool WRAP = MODE == LATTICE_WRAP;
  Sys sys;
  sys.load_params(a.pars, 1, 0);
  const long long start = sys.cell_start(), bound = sys.cell_bound();
  for (long long c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
    if (c < first) continue;
    double k[F];
#pragma unroll
    for (int f = 0; f < F; ++f) k[f] = 0.0;
    long long cg = c;
    if (WRAP) {  // periodic images of cells across the ring's seam are computed too: they are ghost cells
      if (cg < 0) cg += a.n_total; else if (cg >= a.n_total) cg -= a.n_total;
    }
    if (cg < start || cg >= bound) continue;  // not a cell of the user's loop
    sys.cell(xs, k, cg, a.t);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const long long idx = base[f] + c;
      if (HAS_XT) a.xt[idx] = a.x[idx] + a.a_dt * k[f];
      a.acc_out[idx] = a.acc_in[idx] + a.b_dt * k[f];
    }
  }
}

template <bool SHARDED, bool HAS_XT>
ODEINTR_DEV void lattice_stage_body(const odeintr_lattice_args& a) {
  constexpr int F = ODEINTR_NFIELDS;
  long long off[F];
  {
    odeintr::system_type_t<lattice_view_t<LATTICE_PLAIN> > sys;
    sys.field_offsets(off, 0);
  }
  // local position of entry (field f, cell c) is base[f] + c: fields are `pitch` apart, cell c sits at c - c0
  long long base[F];
#pragma unroll
  for (int f = 0; f < F; ++f) base[f] = SHARDED ? (off[f] / a.n_total) * a.pitch - a.c0 : off[f];

  // The launch computes the window [lo, hi) of global cells (on a sharded run it may reach below 0 / above
  // n_total), split into one contiguous chunk per block.  Chunks start at multiples of 32 cells counted from
  // local cell 0 (which the host keeps 128-byte aligned), so every warp reads and writes whole lines.
  long long first = a.lo, last = a.hi;
  if (!SHARDED) {  // one GPU: exactly the user loop's range
    odeintr::system_type_t<lattice_view_t<LATTICE_PLAIN> > sys;
    if (sys.cell_start() > first) first = sys.cell_start();
    if (sys.cell_bound() < last) last = sys.cell_bound();
  }
  if (last <= first) return;
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
  if (c_lo >= c_hi) return;

  if (SHARDED) {
    // A chunk whose cells and stencil neighbours (radius a.halo) stay inside [0, n_total) never wraps: the
    // snippet's global indices then differ from local positions by a constant (one per field), so a plain pointer
    // on a shifted base serves it at full speed.  Only the (at most two) chunks at the ring's seam take the
    // wrapping view.
    const bool interior = c_lo - a.halo >= 0 && c_hi + a.halo <= a.n_total && c_lo >= first;
    if (interior && F == 1) {
      lattice_view_t<LATTICE_PLAIN> v;
      v.p = a.xs - a.c0; v.n_total = a.n_total; v.n_local = a.n; v.c0 = a.c0; v.ghost = a.ghost; v.pitch = a.pitch;
      lattice_chunk<LATTICE_PLAIN, HAS_XT, F>(a, v, base, c_lo, c_hi, first);
      return;
    }
    if (interior) {
      lattice_view_t<LATTICE_FIELDS> v;
      v.p = a.xs; v.n_total = a.n_total; v.n_local = a.n; v.c0 = a.c0; v.ghost = a.ghost; v.pitch = a.pitch;
      lattice_chunk<LATTICE_FIELDS, HAS_XT, F>(a, v, base, c_lo, c_hi, first);
      return;
    }
    lattice_view_t<LATTICE_WRAP> v;
    v.p = a.xs; v.n_total = a.n_total; v.n_local = a.n; v.c0 = a.c0; v.ghost = a.ghost; v.pitch = a.pitch;
    lattice_chunk<LATTICE_WRAP, HAS_XT, F>(a, v, base, c_lo, c_hi, first);
  } else {
    lattice_view_t<LATTICE_PLAIN> v;
    v.p = a.xs; v.n_total = a.n_total; v.n_local = a.n; v.c0 = 0; v.ghost = 0; v.pitch = 0;
#ifdef ODEINTR_LATTICE_2D
    // M x M lattice (the snippet uses laplace2D / M): walk it in column strips of one block width, row after row,
    // so the rows above and below the current one are L1 hits instead of lines 8*M bytes away that come back
    // from DRAM (measured 64 % -> see profiles/ of the HBM peak with the flat order at M = 16384)
    constexpr long long MM = odeintr_b200::isqrt(ODEINTR_SYS_SIZE);
    if (F == 1 && first == 0 && last == MM * MM && MM % 256 == 0 && blockDim.x == 256) {
      typedef odeintr::system_type_t<lattice_view_t<LATTICE_PLAIN> > Sys;
      Sys sys;
      sys.load_params(a.pars, 1, 0);
      const long long strips = MM / 256, total = strips * MM;  // (strip, row) pairs, strip-major
      const long long per_block = (total + gridDim.x - 1) / gridDim.x;
      const long long q_lo = (long long)blockIdx.x * per_block;
      const long long q_hi = q_lo + per_block < total ? q_lo + per_block : total;
      for (long long q = q_lo; q < q_hi; ++q) {
        const long long strip = q / MM, row = q - strip * MM;
        const long long c = row * MM + strip * 256 + threadIdx.x;
        double k[F];
        k[0] = 0.0;
        sys.cell(v, k, c, a.t);
        const long long idx = base[0] + c;
        if (HAS_XT) a.xt[idx] = a.x[idx] + a.a_dt * k[0];
        a.acc_out[idx] = a.acc_in[idx] + a.b_dt * k[0];
      }
      return;
    }
#endif
    lattice_chunk<LATTICE_PLAIN, HAS_XT, F>(a, v, base, c_lo, c_hi, first);
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
