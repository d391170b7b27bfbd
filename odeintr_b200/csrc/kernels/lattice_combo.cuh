// odeintr_b200/csrc/kernels/lattice_combo.cuh
// Generic fused stage kernel for large systems under steppers whose stages cannot share one running sum
// (Dormand-Prince 5(4): every stage has its own coefficient row, the k-vectors stay in HBM):
//     k    = f(xs)[cell]                                   (skipped when xs == null: pure vector algebra)
//     v    = base + c0*in0 + c1*in1 + ... + ck*k           (left to right, like Boost's scale_sumN)
//     out  = v ;  k_out = k
//     err  = max over cells of |v| / (eps_abs + eps_rel * (|err_x| + adt * |err_d|))      (optional)
// With `all` set the same sum is formed for the entries no user loop writes (k = 0 there).
// One launch per Runge-Kutta stage; the error norm of default_error_checker is reduced in the same pass that
// forms the embedded error estimate (warp shuffle max + one atomicMax per warp on the bit pattern, which is
// order preserving for non-negative doubles).
#pragma once
#include "odeintr_b200/lattice_view.cuh"

extern "C" __global__ void __launch_bounds__(256)
odeintr_lattice_combo(const odeintr_lattice_combo_args a) {
  typedef odeintr::system_type_t<odeintr_b200::lattice_view_t<odeintr_b200::LATTICE_PLAIN> > Sys;
  constexpr int F = ODEINTR_NFIELDS;
  Sys sys;
  sys.load_params(a.pars, 1, 0);
  long long off[F];
  sys.field_offsets(off, 0);
  odeintr_b200::lattice_view_t<odeintr_b200::LATTICE_PLAIN> xs;
  xs.p = a.xs; xs.n_total = a.n_total; xs.n_local = a.n_total; xs.c0 = 0; xs.ghost = 0; xs.pitch = 0;

  const long long first = sys.cell_start(), last = sys.cell_bound();
  double emax = 0.0;
  if (last > first) {
    long long chunk = (last - first + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + blockDim.x - 1) / blockDim.x * blockDim.x;
    const long long c_lo = first + (long long)blockIdx.x * chunk;
    long long c_hi = c_lo + chunk;
    if (c_hi > last) c_hi = last;
    for (long long c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
      double k[F];
#pragma unroll
      for (int f = 0; f < F; ++f) k[f] = 0.0;
      if (a.xs) sys.cell(xs, k, c, a.t);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const long long idx = off[f] + c;
        double v;
        int j = 0;
        if (a.base) v = a.base[idx];
        else if (a.nin > 0) { v = a.c[0] * a.in[0][idx]; j = 1; }
        else v = 0.0;
        for (; j < a.nin; ++j) v = v + a.c[j] * a.in[j][idx];
        if (a.xs) {
          if (a.ck != 0.0) v = v + a.ck * k[f];  // a zero weight is a skipped tableau entry, not a term
          if (a.k_out) a.k_out[idx] = k[f];
        }
        if (a.out) a.out[idx] = v;
        if (a.err_max) {
          const double e = fabs(v) / (a.eps_abs + a.eps_rel * (1.0 * fabs(a.err_x[idx]) + a.adt * fabs(a.err_d[idx])));
          emax = fmax(emax, e);
        }
      }
    }
  }
  // the entries no user loop writes (fixed ends, gaps between fields): their derivative is zero
  if (a.all && !(F == 1 && off[0] == 0 && first <= 0 && last >= a.n_total)) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.n_total; idx += (long long)gridDim.x * blockDim.x) {
      bool covered = false;
#pragma unroll
      for (int f = 0; f < F; ++f) covered = covered || (idx - off[f] >= first && idx - off[f] < last);
      if (covered) continue;
      double v;
      int j = 0;
      if (a.base) v = a.base[idx];
      else if (a.nin > 0) { v = a.c[0] * a.in[0][idx]; j = 1; }
      else v = 0.0;
      for (; j < a.nin; ++j) v = v + a.c[j] * a.in[j][idx];
      if (a.xs) {
        if (a.ck != 0.0) v = v + a.ck * 0.0;
        if (a.k_out) a.k_out[idx] = 0.0;
      }
      if (a.out) a.out[idx] = v;
      if (a.err_max) {
        const double e = fabs(v) / (a.eps_abs + a.eps_rel * (1.0 * fabs(a.err_x[idx]) + a.adt * fabs(a.err_d[idx])));
        emax = fmax(emax, e);
      }
    }
  }
  if (a.err_max) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, s));
    if ((threadIdx.x & 31) == 0 && emax > 0.0) atomicMax(a.err_max, (unsigned long long)__double_as_longlong(emax));
  }
}
