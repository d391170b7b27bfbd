// odeintr_b200/csrc/kernels/lattice_dopri5_tiled.cuh
// K4d: one ATTEMPTED Dormand-Prince 5(4) step of a large 1-D system in ONE kernel pass -- the temporally blocked
// counterpart of the seven odeintr_lattice_combo launches behind compile_sys_openmp's default method rk5_i (and
// rk5_a): runge_kutta_dopri5::do_step_impl + default_error_checker::error, i.e.
//     (x, k1 = f(x)) -> x_new, k7 = f(x_new), max-norm of the embedded error estimate
// or, for an observation inside the step, the same stages followed by the dense-output polynomial (k3..k6 never
// leave the registers: interpolating means repeating the accepted attempt, not keeping four more state-sized arrays).
// A block owns a tile of T cells.  Stage states are formed on windows that shrink by `halo` per stage from
// T + 12 halo cells (six right-hand sides depend on neighbours: k2 .. k7), so the attempt reads x and k1 once and
// writes x_new and k7 once: 32 B per cell instead of the ~260 B the stage-per-launch schedule moves.
// Every thread owns CPT cells of the tile and one margin cell; their x and k1 .. k6 live in registers, only the
// stage states y_s travel through two shared-memory windows.  The window is loaded with the cell index wrapped
// periodically, so every tile (seam, ragged end) is treated alike: cells beyond the lattice end are periodic images
// that are computed but not stored.  Operation order and coefficients are those of lattice_adaptive.h (Boost's
// scale_sumN, left to right, coefficient * dt first): bit-identical to the staged path, controller on the host.
#pragma once
#include "odeintr_b200/lattice_tiled.cuh"

#if !defined(ODEINTR_LATTICE_2D) && !defined(ODEINTR_NO_SYMBOLIC) && ODEINTR_LATTICE_STAGES == 7

#ifndef ODEINTR_TILE5_CPT
#define ODEINTR_TILE5_CPT (ODEINTR_NFIELDS > 1 ? 2 : 4)
#endif
#ifndef ODEINTR_TILE5_MIN_BLOCKS
#define ODEINTR_TILE5_MIN_BLOCKS 3
#endif

namespace odeintr_b200 {

// General form: any radius, tiles at the periodic seam, ragged last tile, loops that leave cells alone.
ODEINTR_DEV void lattice_dopri5_tiled_general(const odeintr_lattice_dopri5_args& a) {
  typedef lattice_rel_view<true> View;
  typedef odeintr::system_type_t<View> Sys;
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int C = ODEINTR_TILE5_CPT;
  constexpr int T = 256 * C;
  constexpr long long L = static_cast<long long>(ODEINTR_SYS_SIZE) / ODEINTR_NFIELDS;
  constexpr int NC = C + 1;  // the thread's cells: C of the tile, then its margin cell
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const int h = (int)a.halo, margin = 6 * h, W = T + 2 * margin;
  double* const SA = smem;
  double* const SB = smem + F * W;

  Sys sys;
  sys.load_params(a.pars, 1, 0);
  long long off[F];
  sys.field_offsets(off, 0);
  int fld[F];
#pragma unroll
  for (int f = 0; f < F; ++f) fld[f] = static_cast<int>(off[f] / L);
  const long long start = sys.cell_start(), bound = sys.cell_bound();
  const long long tile_lo = a.own_lo + (long long)blockIdx.x * T;

  // window position, distance from the tile (0 inside), global cell (wrapped), place in the arrays and role of each of
  // the thread's cells
  int jj[NC], dist[NC];
  long long cg[NC], at[NC];
  bool live[NC], store[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    if (k < C) { jj[k] = tid + k * 256; dist[k] = 0; }
    else if (tid < margin) { jj[k] = -1 - tid; dist[k] = tid + 1; }
    else if (tid < 2 * margin) { jj[k] = T + (tid - margin); dist[k] = tid - margin + 1; }
    else { jj[k] = 0; dist[k] = 1 << 20; }  // no margin cell
    long long c = tile_lo + jj[k];
    const bool owned = c < a.own_hi;
    at[k] = c - a.c0;  // a slab holds the periodic images in its ghost cells ...
    c %= L;
    if (c < 0) c += L;
    if (!a.sharded) at[k] = c;  // ... one GPU wraps the index
    cg[k] = c;
    live[k] = dist[k] <= margin && c >= start && c < bound;
    store[k] = k < C && owned && live[k];
  }

  double x[NC][F], k1[NC][F], k2[NC][F], k3[NC][F], k4[NC][F], k5[NC][F], k6[NC][F], y[NC][F];
#pragma unroll
  for (int k = 0; k < NC; ++k)
#pragma unroll
    for (int f = 0; f < F; ++f) {
      x[k][f] = k1[k][f] = 0.0;
      if (dist[k] <= margin) {
        x[k][f] = __ldg(a.x_in + fld[f] * a.pitch + at[k]);
        k1[k][f] = __ldg(a.d_in + fld[f] * a.pitch + at[k]);
      }
    }

  const double dt = a.dt, t = a.t;
  const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
  const double b21 = 1.0 / 5.0;
  const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
  const double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
  const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0, b54 = -212.0 / 729.0;
  const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0, b64 = 49.0 / 176.0,
               b65 = -5103.0 / 18656.0;
  const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
  const double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0, dc4 = c4 - 393.0 / 640.0,
               dc5 = c5 - (-92097.0 / 339200.0), dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;

  // y2 = x + dt*b21*k1 on the whole window; cells the loop does not write show x in every stage
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    if (dist[k] > margin) continue;
#pragma unroll
    for (int f = 0; f < F; ++f) {
      y[k][fld[f]] = live[k] ? x[k][f] + (dt * b21) * k1[k][f] : x[k][f];
      SA[fld[f] * W + margin + jj[k]] = y[k][fld[f]];
      if (!live[k]) SB[fld[f] * W + margin + jj[k]] = x[k][f];
    }
  }
  __syncthreads();

  View view;
  view.n_total = L; view.width = W; view.halo = h; view.bad = 0;
  rel_index idx;
  idx.off = 0; idx.cells = static_cast<int>(L); idx.rel = true;
  double emax = 0.0;

#pragma unroll
  for (int stage = 2; stage <= 7; ++stage) {
    const double* src = (stage & 1) ? SB : SA;
    double* dst = (stage & 1) ? SA : SB;
    const double ts = stage == 2 ? t + dt * a2 : (stage == 3 ? t + dt * a3 : (stage == 4 ? t + dt * a4 :
                      (stage == 5 ? t + dt * a5 : t + dt)));
    const int reach = (7 - stage) * h;  // this stage's window: the tile and `reach` cells on each side
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      if (!live[k] || dist[k] > reach) continue;
      double kk[F];
#pragma unroll
      for (int f = 0; f < F; ++f) kk[f] = 0.0;
      view.p = src + margin + jj[k];
      view.own = y[k];
      view.centre = idx.centre = static_cast<int>(cg[k]);
      sys.cell(view, kk, idx, ts);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const int q = fld[f] * W + margin + jj[k];
        const long long g = fld[f] * a.pitch + at[k];
        double v;
        if (stage == 2) {
          k2[k][f] = kk[f];
          v = x[k][f] + (dt * b31) * k1[k][f];
          v = v + (dt * b32) * kk[f];
        } else if (stage == 3) {
          k3[k][f] = kk[f];
          v = x[k][f] + (dt * b41) * k1[k][f];
          v = v + (dt * b42) * k2[k][f];
          v = v + (dt * b43) * kk[f];
        } else if (stage == 4) {
          k4[k][f] = kk[f];
          v = x[k][f] + (dt * b51) * k1[k][f];
          v = v + (dt * b52) * k2[k][f];
          v = v + (dt * b53) * k3[k][f];
          v = v + (dt * b54) * kk[f];
        } else if (stage == 5) {
          k5[k][f] = kk[f];
          v = x[k][f] + (dt * b61) * k1[k][f];
          v = v + (dt * b62) * k2[k][f];
          v = v + (dt * b63) * k3[k][f];
          v = v + (dt * b64) * k4[k][f];
          v = v + (dt * b65) * kk[f];
        } else if (stage == 6) {
          k6[k][f] = kk[f];
          v = x[k][f] + (dt * c1) * k1[k][f];
          v = v + (dt * c3) * k3[k][f];
          v = v + (dt * c4) * k4[k][f];
          v = v + (dt * c5) * k5[k][f];
          v = v + (dt * c6) * kk[f];
          if (store[k] && !a.dense_out) a.x_out[g] = v;  // x_new
        } else {
          // k7 = f(x_new): the derivative handed to the next step, and the last term of the error estimate
          if (a.dense_out) {
            if (store[k]) {
              double w = x[k][f] + a.dcoef[0] * k1[k][f];
              w = w + a.dcoef[1] * k3[k][f];
              w = w + a.dcoef[2] * k4[k][f];
              w = w + a.dcoef[3] * k5[k][f];
              w = w + a.dcoef[4] * k6[k][f];
              w = w + a.dcoef[5] * kk[f];
              a.dense_out[g] = w;
            }
            v = 0.0;
          } else {
            if (store[k]) a.d_out[g] = kk[f];
            v = (dt * dc1) * k1[k][f];
            v = v + (dt * dc3) * k3[k][f];
            v = v + (dt * dc4) * k4[k][f];
            v = v + (dt * dc5) * k5[k][f];
            v = v + (dt * dc6) * k6[k][f];
            v = v + (dt * dc7) * kk[f];
            if (store[k]) {
              const double e = fabs(v) / (a.eps_abs + a.eps_rel * (1.0 * fabs(x[k][f]) + a.adt * fabs(k1[k][f])));
              emax = fmax(emax, e);
            }
          }
        }
        if (stage < 7) {
          dst[q] = v;
          y[k][fld[f]] = v;
        }
      }
    }
    if (stage < 7) __syncthreads();
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, s));
  if ((tid & 31) == 0 && emax > 0.0) atomicMax(a.err_max, (unsigned long long)__double_as_longlong(emax));
  if (view.bad) *a.flag = 1;
}

// Fast form for a compile-time radius H: tiles that are full, away from the periodic seam and made of cells the
// loop writes need no per-cell predicates, no index wrapping and only immediate shared-memory offsets.  Everything
// else in the same launch takes the general form above (same arithmetic).
template <int H>
ODEINTR_DEV void lattice_dopri5_tiled_fast(const odeintr_lattice_dopri5_args& a) {
  typedef lattice_rel_view<true> View;
  typedef odeintr::system_type_t<View> Sys;
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int C = ODEINTR_TILE5_CPT;
  constexpr int T = 256 * C;
  constexpr long long L = static_cast<long long>(ODEINTR_SYS_SIZE) / ODEINTR_NFIELDS;
  constexpr int margin = 6 * H, W = T + 2 * margin;
  Sys sys;
  const long long start = sys.cell_start(), bound = sys.cell_bound();
  const long long tile_lo = a.own_lo + (long long)blockIdx.x * T;
  // block-uniform: a full tile whose whole window lies inside the lattice and inside the user loop's range
  if (!(tile_lo - margin >= (start > 0 ? start : 0) && tile_lo + T + margin <= (bound < L ? bound : L) &&
        tile_lo + T <= a.own_hi)) {
    lattice_dopri5_tiled_general(a);
    return;
  }
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  double* const SA = smem;
  double* const SB = smem + F * W;
  sys.load_params(a.pars, 1, 0);
  long long off[F];
  sys.field_offsets(off, 0);
  int fld[F];
#pragma unroll
  for (int f = 0; f < F; ++f) fld[f] = static_cast<int>(off[f] / L);

  // the thread's margin cell: window position jm at distance dm from the tile (dm > margin: none)
  const bool has_m = tid < 2 * margin;
  const int jm = tid < margin ? -1 - tid : T + (tid - margin);
  const int dm = has_m ? (tid < margin ? tid + 1 : tid - margin + 1) : (1 << 20);

  double x[C][F], k1[C][F], k2[C][F], k3[C][F], k4[C][F], k5[C][F], k6[C][F], y[C][F];
  double xm[F], m1[F], m2[F], m3[F], m4[F], m5[F], ym[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const double* gx = a.x_in + fld[f] * a.pitch + (tile_lo - a.c0);
    const double* gd = a.d_in + fld[f] * a.pitch + (tile_lo - a.c0);
#pragma unroll
    for (int k = 0; k < C; ++k) {
      x[k][f] = __ldg(gx + tid + k * 256);
      k1[k][f] = __ldg(gd + tid + k * 256);
    }
    xm[f] = has_m ? __ldg(gx + jm) : 0.0;
    m1[f] = has_m ? __ldg(gd + jm) : 0.0;
  }

  const double dt = a.dt, t = a.t;
  const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
  const double b21 = 1.0 / 5.0;
  const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
  const double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
  const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0, b54 = -212.0 / 729.0;
  const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0, b64 = 49.0 / 176.0,
               b65 = -5103.0 / 18656.0;
  const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
  const double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0, dc4 = c4 - 393.0 / 640.0,
               dc5 = c5 - (-92097.0 / 339200.0), dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;

  // y2 = x + dt*b21*k1 on the whole window
#pragma unroll
  for (int f = 0; f < F; ++f) {
#pragma unroll
    for (int k = 0; k < C; ++k) {
      y[k][fld[f]] = x[k][f] + (dt * b21) * k1[k][f];
      SA[fld[f] * W + margin + tid + k * 256] = y[k][fld[f]];
    }
    if (has_m) {
      ym[fld[f]] = xm[f] + (dt * b21) * m1[f];
      SA[fld[f] * W + margin + jm] = ym[fld[f]];
    }
  }
  __syncthreads();

  View view;
  view.n_total = L; view.width = W; view.halo = H; view.bad = 0;
  rel_index idx;
  idx.off = 0; idx.cells = static_cast<int>(L); idx.rel = true;
  double emax = 0.0;

#pragma unroll
  for (int stage = 2; stage <= 7; ++stage) {
    const double* src = (stage & 1) ? SB : SA;
    double* dst = (stage & 1) ? SA : SB;
    const double ts = stage == 2 ? t + dt * a2 : (stage == 3 ? t + dt * a3 : (stage == 4 ? t + dt * a4 :
                      (stage == 5 ? t + dt * a5 : t + dt)));
    // the stage's linear combination, in Boost's order; kk = this stage's derivative
    auto combine = [&](const int f, const double X, const double K1, const double K2, const double K3, const double K4,
                       const double K5, const double K6, const double kk) {
      double v;
      if (stage == 2) {
        v = X + (dt * b31) * K1;
        v = v + (dt * b32) * kk;
      } else if (stage == 3) {
        v = X + (dt * b41) * K1;
        v = v + (dt * b42) * K2;
        v = v + (dt * b43) * kk;
      } else if (stage == 4) {
        v = X + (dt * b51) * K1;
        v = v + (dt * b52) * K2;
        v = v + (dt * b53) * K3;
        v = v + (dt * b54) * kk;
      } else if (stage == 5) {
        v = X + (dt * b61) * K1;
        v = v + (dt * b62) * K2;
        v = v + (dt * b63) * K3;
        v = v + (dt * b64) * K4;
        v = v + (dt * b65) * kk;
      } else if (stage == 6) {
        v = X + (dt * c1) * K1;
        v = v + (dt * c3) * K3;
        v = v + (dt * c4) * K4;
        v = v + (dt * c5) * K5;
        v = v + (dt * c6) * kk;
      } else {
        v = (dt * dc1) * K1;
        v = v + (dt * dc3) * K3;
        v = v + (dt * dc4) * K4;
        v = v + (dt * dc5) * K5;
        v = v + (dt * dc6) * K6;
        v = v + (dt * dc7) * kk;
      }
      (void)f;
      return v;
    };
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const int j = tid + k * 256;
      double kk[F];
#pragma unroll
      for (int f = 0; f < F; ++f) kk[f] = 0.0;
      view.p = src + margin + j;
      view.own = y[k];
      view.centre = idx.centre = static_cast<int>(tile_lo) + j;
      sys.cell(view, kk, idx, ts);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const long long g = fld[f] * a.pitch + (tile_lo - a.c0) + j;
        const double v = combine(f, x[k][f], k1[k][f], k2[k][f], k3[k][f], k4[k][f], k5[k][f], k6[k][f], kk[f]);
        if (stage == 2) k2[k][f] = kk[f];
        if (stage == 3) k3[k][f] = kk[f];
        if (stage == 4) k4[k][f] = kk[f];
        if (stage == 5) k5[k][f] = kk[f];
        if (stage == 6) { k6[k][f] = kk[f]; if (!a.dense_out) a.x_out[g] = v; }
        if (stage == 7) {
          if (a.dense_out) {
            double w = x[k][f] + a.dcoef[0] * k1[k][f];
            w = w + a.dcoef[1] * k3[k][f];
            w = w + a.dcoef[2] * k4[k][f];
            w = w + a.dcoef[3] * k5[k][f];
            w = w + a.dcoef[4] * k6[k][f];
            w = w + a.dcoef[5] * kk[f];
            a.dense_out[g] = w;
          } else {
            a.d_out[g] = kk[f];
            const double e = fabs(v) / (a.eps_abs + a.eps_rel * (1.0 * fabs(x[k][f]) + a.adt * fabs(k1[k][f])));
            emax = fmax(emax, e);
          }
        } else {
          dst[fld[f] * W + margin + j] = v;
          y[k][fld[f]] = v;
        }
      }
    }
    if (stage < 7) {
      if (dm <= (7 - stage) * H) {
        double kk[F];
#pragma unroll
        for (int f = 0; f < F; ++f) kk[f] = 0.0;
        view.p = src + margin + jm;
        view.own = ym;
        view.centre = idx.centre = static_cast<int>(tile_lo) + jm;
        sys.cell(view, kk, idx, ts);
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const double v = combine(f, xm[f], m1[f], m2[f], m3[f], m4[f], m5[f], 0.0, kk[f]);
          if (stage == 2) m2[f] = kk[f];
          if (stage == 3) m3[f] = kk[f];
          if (stage == 4) m4[f] = kk[f];
          if (stage == 5) m5[f] = kk[f];
          dst[fld[f] * W + margin + jm] = v;
          ym[fld[f]] = v;
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, s));
  if ((tid & 31) == 0 && emax > 0.0) atomicMax(a.err_max, (unsigned long long)__double_as_longlong(emax));
  if (view.bad) *a.flag = 1;
}

}  // namespace odeintr_b200

extern "C" __global__ void __launch_bounds__(256, ODEINTR_TILE5_MIN_BLOCKS)
odeintr_lattice_dopri5_tiled(const odeintr_lattice_dopri5_args a) { odeintr_b200::lattice_dopri5_tiled_general(a); }

extern "C" __global__ void __launch_bounds__(256, ODEINTR_TILE5_MIN_BLOCKS)
odeintr_lattice_dopri5_tiled_h1(const odeintr_lattice_dopri5_args a) { odeintr_b200::lattice_dopri5_tiled_fast<1>(a); }

extern "C" __global__ void __launch_bounds__(256, ODEINTR_TILE5_MIN_BLOCKS)
odeintr_lattice_dopri5_tiled_h2(const odeintr_lattice_dopri5_args a) { odeintr_b200::lattice_dopri5_tiled_fast<2>(a); }

#endif
