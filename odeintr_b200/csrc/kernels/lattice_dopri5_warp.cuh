// odeintr_b200/csrc/kernels/lattice_dopri5_warp.cuh
// K4dw: one ATTEMPTED Dormand-Prince 5(4) step of a large 1-D system (lattice_dopri5_tiled.cuh, compile_sys_openmp's
// default method rk5_i and rk5_a) with the warp-autonomous scheme of lattice_warp.cuh:
// a warp owns a window of 32 * C consecutive cells, every lane C consecutive cells of it -- their x, k1 .. k6 and the
// current stage state live in the lane's registers for the whole attempt; the H cells beyond a lane's ends come from
// the neighbouring lanes by shuffles; no stage state touches shared memory and no barrier separates the six
// neighbour-dependent stages (k2 .. k7).  Every lane evaluates all its cells in every stage; what the window's ends
// lack spoils H more cells per stage, so the outer 6 * H cells on either side are not stored: a warp tile yields
// P = 32 * C - 12 * H cells (C = 5, H = 1: 148 of 160).  x and k1 arrive through the warp's own double-buffered TMA
// pipeline (two tiles ahead), x_new and k7 (or the dense-output row) leave through staging buffers and TMA bulk stores;
// the error norm is reduced by shuffles and one atomicMax per warp.
// Operation order and coefficients are those of lattice_dopri5_tiled.cuh / lattice_adaptive.h (Boost's scale_sumN, left
// to right, coefficient * dt first): bit-identical to them; the controller stays on the host.
// The launch produces the cells [own_lo, own_hi) (the host hands it the interior of the lattice or slab: windows inside
// the user loop's range, away from the periodic seam); everything else stays with lattice_dopri5_tiled.cuh.
#pragma once
#include "odeintr_b200/lattice_warp.cuh"

#if !defined(ODEINTR_LATTICE_2D) && !defined(ODEINTR_NO_SYMBOLIC) && ODEINTR_LATTICE_STAGES == 7

#ifndef ODEINTR_WARP5_CPT  // cells per lane: odd, >= the radius
#define ODEINTR_WARP5_CPT (ODEINTR_NFIELDS > 1 ? 3 : 7)  // measured (profiles/r02_sweep_dopri5_warp.log): 7 cells / 2 blocks per SM best
#endif
#ifndef ODEINTR_WARP5_MIN_BLOCKS
#define ODEINTR_WARP5_MIN_BLOCKS 2
#endif

namespace odeintr_b200 {

template <int H>
ODEINTR_DEV void lattice_dopri5_warp_body(const odeintr_lattice_dopri5_args& a) {
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int C = ODEINTR_WARP5_CPT;
  constexpr int WC = 32 * C;
  constexpr int M6 = 6 * H;
  constexpr int P = WC - 2 * M6;
  constexpr int S = C + 2 * H;
  constexpr int WPB = 8;
  constexpr long long L = static_cast<long long>(ODEINTR_SYS_SIZE) / ODEINTR_NFIELDS;
  static_assert(C % 2 == 1 && C >= H && P > 0, "cells per lane: odd, at least the stencil radius, wider than the margins");
  typedef odeintr::system_type_t<lattice_reg_view<H, C> > Sys;
  extern __shared__ __align__(128) double smem_w[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // per warp: input windows [buffer][x | k1][field][WC], output staging [buffer][x_new | k7][field][WC]
  double* in0 = smem_w + warp * (8 * F * WC);
  double* out0 = in0 + 4 * F * WC;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_w + WPB * 8 * F * WC) + 2 * warp;

  Sys sys;
  sys.load_params(a.pars, 1, 0);
  int fld[F];
  {
    long long off[F];
    sys.field_offsets(off, 0);
#pragma unroll
    for (int f = 0; f < F; ++f) fld[f] = static_cast<int>(off[f] / L);
  }
  const bool tma = a.tma != 0;
  const bool dense = a.dense_out != nullptr;
  if (tma && lane == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const long long n_tiles = (a.own_hi - a.own_lo + P - 1) / P;
  const long long first = static_cast<long long>(blockIdx.x) * WPB + warp, stride = static_cast<long long>(gridDim.x) * WPB;
  auto tile_start = [&](const long long w) {
    const long long s = a.own_lo + w * P;
    return s + P > a.own_hi ? a.own_hi - P : s;  // the last tile is moved back (shared cells: same bits twice)
  };
  const unsigned bytes = static_cast<unsigned>(WC * sizeof(double));
  auto prefetch = [&](const long long w, const int b) {
    const long long ws = tile_start(w) - M6;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar + b, bytes * 2 * F);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      tma_load_1d(in0 + ((b * 2 + 0) * F + fld[f]) * WC, a.x_in + fld[f] * a.pitch + (ws - a.c0), bytes, bar + b);
      tma_load_1d(in0 + ((b * 2 + 1) * F + fld[f]) * WC, a.d_in + fld[f] * a.pitch + (ws - a.c0), bytes, bar + b);
    }
  };
  if (tma && lane == 0) {
    if (first < n_tiles) prefetch(first, 0);
    if (first + stride < n_tiles) prefetch(first + stride, 1);
  }

  lattice_reg_view<H, C> view;
  view.n_total = L; view.bad = 0;
  rel_index idx;
  idx.off = 0; idx.cells = static_cast<int>(L); idx.rel = true;
  const unsigned full = 0xffffffffu;

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
  double emax = 0.0;

  int it = 0;
  for (long long w = first; w < n_tiles; w += stride, ++it) {
    const int b = it & 1;
    const long long s = tile_start(w), ws = s - M6;
    double* inx = in0 + (b * 2 + 0) * F * WC;
    double* ind = in0 + (b * 2 + 1) * F * WC;
    if (tma) {
      mbar_wait(bar + b, (it >> 1) & 1);
    } else {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const double* gx = a.x_in + fld[f] * a.pitch + (ws - a.c0);
        const double* gd = a.d_in + fld[f] * a.pitch + (ws - a.c0);
        double vx[C], vd[C];
#pragma unroll
        for (int k = 0; k < C; ++k) { vx[k] = __ldg(gx + k * 32 + lane); vd[k] = __ldg(gd + k * 32 + lane); }
#pragma unroll
        for (int k = 0; k < C; ++k) { inx[fld[f] * WC + k * 32 + lane] = vx[k]; ind[fld[f] * WC + k * 32 + lane] = vd[k]; }
      }
      __syncwarp();
    }
    double x[C][F], k1[C][F], k2[C][F], k3[C][F], k4[C][F], k5[C][F], k6[C][F], y[C][F];
#pragma unroll
    for (int f = 0; f < F; ++f)
#pragma unroll
      for (int k = 0; k < C; ++k) {
        x[k][f] = inx[fld[f] * WC + lane * C + k];
        k1[k][f] = ind[fld[f] * WC + lane * C + k];
        y[k][fld[f]] = x[k][f] + (dt * b21) * k1[k][f];  // stage state of k2, pointwise on the whole window
      }
    __syncwarp();
    if (tma && lane == 0 && w + 2 * stride < n_tiles) prefetch(w + 2 * stride, b);

    double* outx = out0 + (b * 2 + 0) * F * WC;
    double* outd = out0 + (b * 2 + 1) * F * WC;
#pragma unroll
    for (int stage = 2; stage <= 7; ++stage) {
      const double ts = stage == 2 ? t + dt * a2 : (stage == 3 ? t + dt * a3 : (stage == 4 ? t + dt * a4 :
                        (stage == 5 ? t + dt * a5 : t + dt)));
      double ext[F][S];
#pragma unroll
      for (int fn = 0; fn < F; ++fn) {
#pragma unroll
        for (int k = 0; k < C; ++k) ext[fn][H + k] = y[k][fn];
#pragma unroll
        for (int q = 0; q < H; ++q) {
          ext[fn][q] = __shfl_up_sync(full, y[C - H + q][fn], 1);
          ext[fn][H + C + q] = __shfl_down_sync(full, y[q][fn], 1);
        }
      }
      if (stage == 7) {
        // before the staging buffer is written: the store issued from it two tiles ago has read it
        if (tma) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
      }
#pragma unroll
      for (int k = 0; k < C; ++k) {
        double kk[F];
#pragma unroll
        for (int f = 0; f < F; ++f) kk[f] = 0.0;
        view.e = &ext[0][H + k];
        view.centre = idx.centre = static_cast<int>(ws) + lane * C + k;
        sys.cell(view, kk, idx, ts);
        const int j = lane * C + k;
        const bool kept = j >= M6 && j < WC - M6;
#pragma unroll
        for (int f = 0; f < F; ++f) {
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
            v = v + (dt * c6) * kk[f];  // x_new
          } else if (dense) {
            // dense output of the (already accepted) step, the interpolation folded into the last stage
            v = x[k][f] + a.dcoef[0] * k1[k][f];
            v = v + a.dcoef[1] * k3[k][f];
            v = v + a.dcoef[2] * k4[k][f];
            v = v + a.dcoef[3] * k5[k][f];
            v = v + a.dcoef[4] * k6[k][f];
            v = v + a.dcoef[5] * kk[f];
            outx[fld[f] * WC + j] = v;
          } else {
            // k7 = f(x_new): the derivative handed to the next step, and the last term of the error estimate
            outx[fld[f] * WC + j] = y[k][fld[f]];
            outd[fld[f] * WC + j] = kk[f];
            v = (dt * dc1) * k1[k][f];
            v = v + (dt * dc3) * k3[k][f];
            v = v + (dt * dc4) * k4[k][f];
            v = v + (dt * dc5) * k5[k][f];
            v = v + (dt * dc6) * k6[k][f];
            v = v + (dt * dc7) * kk[f];
            if (kept) {
              const double e = fabs(v) / (a.eps_abs + a.eps_rel * (1.0 * fabs(x[k][f]) + a.adt * fabs(k1[k][f])));
              emax = fmax(emax, e);
            }
          }
          if (stage < 7) y[k][fld[f]] = v;
        }
      }
    }

    // staging buffers -> HBM
    const int n_out = dense ? 1 : 2;
    if (tma) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
          double* gx = (dense ? a.dense_out : a.x_out) + fld[f] * a.pitch + (s - a.c0);
          tma_store_1d(gx, outx + fld[f] * WC + M6, static_cast<unsigned>(P * sizeof(double)));
          if (!dense)
            tma_store_1d(a.d_out + fld[f] * a.pitch + (s - a.c0), outd + fld[f] * WC + M6, static_cast<unsigned>(P * sizeof(double)));
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      __syncwarp();
#pragma unroll
      for (int f = 0; f < F; ++f) {
        double* gx = (dense ? a.dense_out : a.x_out) + fld[f] * a.pitch + (s - a.c0);
        for (int j = lane; j < P; j += 32) gx[j] = outx[fld[f] * WC + M6 + j];
        if (!dense) {
          double* gd = a.d_out + fld[f] * a.pitch + (s - a.c0);
          for (int j = lane; j < P; j += 32) gd[j] = outd[fld[f] * WC + M6 + j];
        }
      }
      __syncwarp();
    }
    (void)n_out;
  }
  if (tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) emax = fmax(emax, __shfl_xor_sync(full, emax, sft));
  if (lane == 0 && emax > 0.0) atomicMax(a.err_max, (unsigned long long)__double_as_longlong(emax));
  if (view.bad) *a.flag = 1;
}

}  // namespace odeintr_b200

extern "C" __global__ void __launch_bounds__(256, ODEINTR_WARP5_MIN_BLOCKS)
odeintr_lattice_dopri5_warp_h1(const odeintr_lattice_dopri5_args a) { odeintr_b200::lattice_dopri5_warp_body<1>(a); }

#if ODEINTR_WARP5_CPT >= 2
extern "C" __global__ void __launch_bounds__(256, ODEINTR_WARP5_MIN_BLOCKS)
odeintr_lattice_dopri5_warp_h2(const odeintr_lattice_dopri5_args a) { odeintr_b200::lattice_dopri5_warp_body<2>(a); }
#endif

extern "C" __global__ void odeintr_lattice_dopri5_warp_info(int* out) { out[0] = ODEINTR_WARP5_CPT; }

#endif  // dopri5 on a 1-D lattice
