// odeintr_b200/csrc/kernels/lattice_warp.cuh
// K4w: the temporally blocked rk4 step of a large 1-D system (lattice_tiled.cuh) with WARP-AUTONOMOUS tiles.
//
// K4t gives a block one tile, lets every thread own cells 256 apart and passes the stage states between threads
// through shared memory: four block barriers per tile and a dependent LDS -> DADD/DMUL chain after each of them; ncu
// showed it latency / phase bound at 0.61 of the FP64 issue roofline (profiles/r01_lattice_tiled.md).  Here a WARP owns
// a window of 32 * C consecutive cells and every lane C CONSECUTIVE cells of it, for the whole step, in registers:
//   * the neighbours of a lane's inner cells are the lane's own registers; the H cells beyond its ends come from the
//     neighbouring lanes by __shfl_up / __shfl_down -- 4 * H shuffles per field, lane and stage;
//   * no stage state ever touches shared memory, no barrier separates the stages (the shuffles are the only
//     synchronisation), warps never wait for each other;
//   * nothing is special about the window's ends: every lane computes all its C cells in every stage; what the first
//     and last lane miss (their outer neighbours) makes the outermost H cells of the window wrong after stage 1, 2 H
//     after stage 2, ...: after four stages the outer 4 * H cells on each side are garbage and are simply not stored.
//     A warp tile therefore yields P = 32 * C - 8 * H cells of the new state (C = 7, H = 1: 216 of 224, 3.6 % redundant
//     work -- less than K4t's margins).
//   * HBM traffic stays 8 B in + 8 B out per cell and step.  Each warp runs its own double-buffered TMA pipeline:
//     lane 0 asks the TMA engine for the window two tiles ahead (cp.async.bulk global -> shared, completion counted on
//     the warp's own mbarrier), lanes pick their C consecutive cells out of shared memory (C odd: the 8 * C byte lane
//     stride is bank-conflict free), results go back through a staging buffer and a TMA bulk store
//     (cp.async.bulk shared -> global), so no warp ever waits for HBM and stores are full 128-byte lines.
//     Windows that are not 16-byte aligned use plain coalesced loads / stores through the same buffers (a.tma == 0).
// Every cell goes through exactly the operations of the staged schedule (Boost's order, no FMA): bit-identical.
//
// The radius H is a template parameter (1 and 2 are instantiated; wider stencils stay on K4t): a lane's neighbourhood
// must be a compile-time register window.  Every access is still checked against it; a violation raises the flag and
// the host repeats the call with the stage-per-launch kernels.
#pragma once
#include "odeintr_b200/lattice_tiled.cuh"

// Cells per lane C: odd (bank-conflict-free lane stride), >= the radius.  Measured on a B200 (2^28 states, profiles/
// r02_lattice_warp.md): more cells per lane amortise the per-tile work of lane 0 (TMA issue) and shrink the margins'
// share, and with one block of 8 warps per SM the compiler may keep all C independent cells in flight (no register
// cap): one field C = 7 / 3 blocks 1.00 ms per step, 11 / 2 blocks 0.95, 13 / 1 block 0.90; two fields 5 / 2 blocks
// 0.81, 9 / 1 block 0.74.
#ifndef ODEINTR_WARP_CPT
#define ODEINTR_WARP_CPT (ODEINTR_NFIELDS > 2 ? 5 : (ODEINTR_NFIELDS > 1 ? 9 : 13))
#endif
#ifndef ODEINTR_WARP_MIN_BLOCKS
#define ODEINTR_WARP_MIN_BLOCKS 1
#endif

#if !defined(ODEINTR_LATTICE_2D) && !defined(ODEINTR_NO_SYMBOLIC)

namespace odeintr_b200 {

// View of one stage's input for cell k of a lane: `e` points at (field 0, cell k) inside the lane's register window
// ext[field][C + 2 H] (H neighbour cells on each side of its C own cells).  Field and displacement of every access
// of an ordinary stencil are compile-time constants after inlining, so the select chain below folds to one register.
template <int H, int C>
struct lattice_reg_view {
  static constexpr int S = C + 2 * H;  // doubles between fields
  const double* e;
  long long n_total;  // cells per field
  int centre;         // global number of the cell being computed (absolute-index path)
  mutable int bad;    // a read reached beyond the radius or outside the state

  ODEINTR_DEV double at(long long f, long long d) const {
    // out-of-reach reads are served from a clamped position and reported once per thread at the end of the kernel
    if (f < 0) { bad = 1; f = 0; }
    if (f >= ODEINTR_NFIELDS) { bad = 1; f = ODEINTR_NFIELDS - 1; }
    if (d > H) { bad = 1; d = H; }
    if (d < -H) { bad = 1; d = -H; }
    double v = 0.0;
#pragma unroll
    for (int ff = 0; ff < ODEINTR_NFIELDS; ++ff)
#pragma unroll
      for (int dd = -H; dd <= H; ++dd)
        if (f == ff && d == dd) v = e[ff * S + dd];  // never a run-time register index
    return v;
  }
  ODEINTR_DEV double operator[](const rel_index& r) const {
    if (!r.rel) return (*this)[r.off];
    const long long L = n_total;  // off = f * L + d with |d| small: nearest multiple of L
    const long long f = r.off >= 0 ? (r.off + L / 2) / L : -((-r.off + L / 2) / L);
    return at(f, r.off - f * L);
  }
  // absolute index: no periodic image is within reach of an interior cell, the distance is taken as is
  ODEINTR_DEV double operator[](const long long g) const {
    const long long f = g >= 0 ? g / n_total : -1;
    return at(f, g - f * n_total - centre);
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(n_total); }
};

ODEINTR_DEV void tma_store_1d(void* gdst, const void* ssrc, const unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)), "r"(bytes)
               : "memory");
}

}  // namespace odeintr_b200

#endif

#if !defined(ODEINTR_LATTICE_2D) && !defined(ODEINTR_NO_SYMBOLIC) && ODEINTR_LATTICE_STAGES == 4

namespace odeintr_b200 {

// Shared memory of a block of 8 warps: per warp two input windows and two output staging windows of F * 32 C doubles,
// then two mbarriers per warp.
template <int H>
ODEINTR_DEV void lattice_rk4_warp_body(const odeintr_lattice_tiled_args& a) {
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int C = ODEINTR_WARP_CPT;
  constexpr int WC = 32 * C;      // cells of a warp's window
  // Steps per pass: a.nsub = 2 takes TWO rk4 steps (at a.t and a.t2) before anything goes back to HBM -- 8 bytes per
  // cell and step instead of 16 for systems whose single step is bound by HBM (the two-field oscillator chain), at
  // the price of a margin twice as wide.
  const int nsub = a.nsub > 1 ? 2 : 1;
  const int M4 = 4 * H * nsub;    // margin: cells on either side of the window that do not survive the pass
  const int P = WC - 2 * M4;      // cells of the new state a warp tile yields
  constexpr int S = C + 2 * H;
  constexpr int WPB = 8;          // warps per block
  constexpr long long L = static_cast<long long>(ODEINTR_SYS_SIZE) / ODEINTR_NFIELDS;
  static_assert(C % 2 == 1 && C >= H, "cells per lane: odd and at least the stencil radius");
  typedef odeintr::system_type_t<lattice_reg_view<H, C> > Sys;
  extern __shared__ __align__(128) double smem_w[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* in0 = smem_w + warp * (4 * F * WC);
  double* out0 = in0 + 2 * F * WC;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_w + WPB * 4 * F * WC) + 2 * warp;

  Sys sys;
  sys.load_params(a.pars, 1, 0);
  int fld[F];  // field number of the f-th written entry
  {
    long long off[F];
    sys.field_offsets(off, 0);
#pragma unroll
    for (int f = 0; f < F; ++f) fld[f] = static_cast<int>(off[f] / L);
  }
  const bool tma = a.tma != 0;
  if (tma && lane == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // warp tile w yields the cells [s, s + P), s = lo + w * P; the last one is moved back to end at hi (the cells it
  // shares with its neighbour are written twice with the same bits)
  const long long n_tiles = (a.hi - a.lo + P - 1) / P;
  const long long first = static_cast<long long>(blockIdx.x) * WPB + warp, stride = static_cast<long long>(gridDim.x) * WPB;
  auto tile_start = [&](const long long w) {
    const long long s = a.lo + w * P;
    return s + P > a.hi ? a.hi - P : s;
  };
  const unsigned bytes = static_cast<unsigned>(WC * sizeof(double));
  auto prefetch = [&](const long long w, const int b) {  // lane 0: window of tile w -> input buffer b
    const long long ws = tile_start(w) - M4;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the lanes' reads of this buffer come first
    mbar_expect_tx(bar + b, bytes * F);
#pragma unroll
    for (int f = 0; f < F; ++f)
      tma_load_1d(in0 + (b * F + fld[f]) * WC, a.x_in + fld[f] * a.pitch + (ws - a.c0), bytes, bar + b);
  };
  if (tma && lane == 0) {
    if (first < n_tiles) prefetch(first, 0);
    if (first + stride < n_tiles) prefetch(first + stride, 1);
  }

  lattice_reg_view<H, C> view;
  view.n_total = L; view.bad = 0;
  rel_index idx;
  idx.off = 0; idx.cells = static_cast<int>(L); idx.rel = true;
  const double half = 1.0 / 2.0, b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
  const double dt = a.dt;
  const unsigned full = 0xffffffffu;

  int it = 0;
  for (long long w = first; w < n_tiles; w += stride, ++it) {
    const int b = it & 1;
    const long long s = tile_start(w), ws = s - M4;
    double* in = in0 + b * F * WC;
    if (tma) {
      mbar_wait(bar + b, (it >> 1) & 1);
    } else {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const double* g = a.x_in + fld[f] * a.pitch + (ws - a.c0);
        double v[C];  // all loads in flight before the first store
#pragma unroll
        for (int k = 0; k < C; ++k) v[k] = __ldg(g + k * 32 + lane);
#pragma unroll
        for (int k = 0; k < C; ++k) in[fld[f] * WC + k * 32 + lane] = v[k];
      }
      __syncwarp();
    }
    // the lane's C consecutive cells: x at the start of the step (by written entry) and the stage input (by field)
    double xs[C][F], acc[C][F], own[C][F];
#pragma unroll
    for (int f = 0; f < F; ++f)
#pragma unroll
      for (int k = 0; k < C; ++k) {
        xs[k][f] = in[fld[f] * WC + lane * C + k];
        own[k][fld[f]] = xs[k][f];
      }
    __syncwarp();  // every lane has its cells: the input buffer is free for the tile after the next
    if (tma && lane == 0 && w + 2 * stride < n_tiles) prefetch(w + 2 * stride, b);

#pragma unroll 1
    for (int sub = 0; sub < nsub; ++sub) {
    const double t_step = sub == 0 ? a.t : a.t2;
#pragma unroll
    for (int stage = 1; stage <= 4; ++stage) {
      const double a_dt = stage == 1 ? half * dt : (stage == 2 ? half * dt : 1.0 * dt);
      const double b_dt = stage == 1 ? b1 * dt : (stage == 2 ? b2 * dt : (stage == 3 ? b2 * dt : b1 * dt));
      const double ts = stage == 1 ? t_step : (stage == 4 ? t_step + 1.0 * dt : t_step + half * dt);
      // register window of the stage input: own cells + H cells of either neighbour lane
      double ext[F][S];
#pragma unroll
      for (int fn = 0; fn < F; ++fn) {
#pragma unroll
        for (int k = 0; k < C; ++k) ext[fn][H + k] = own[k][fn];
#pragma unroll
        for (int q = 0; q < H; ++q) {
          ext[fn][q] = __shfl_up_sync(full, own[C - H + q][fn], 1);
          ext[fn][H + C + q] = __shfl_down_sync(full, own[q][fn], 1);
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
#pragma unroll
        for (int f = 0; f < F; ++f) {
          if (stage < 4) own[k][fld[f]] = xs[k][f] + a_dt * kk[f];
          acc[k][f] = (stage == 1 ? xs[k][f] : acc[k][f]) + b_dt * kk[f];  // the running sum starts from x
        }
      }
    }
    if (sub + 1 < nsub) {  // the state after the first step is the start of the second
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int k = 0; k < C; ++k) {
          xs[k][f] = acc[k][f];
          own[k][fld[f]] = acc[k][f];
        }
    }
    }

    // new state -> staging buffer -> HBM (the window's outer 4 H cells per step on each side are not part of it)
    double* out = out0 + b * F * WC;
    if (tma) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store issued from `out` two tiles ago
      __syncwarp();
    }
#pragma unroll
    for (int f = 0; f < F; ++f)
#pragma unroll
      for (int k = 0; k < C; ++k) out[fld[f] * WC + lane * C + k] = acc[k][f];
    if (tma) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the lanes' stores become visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int f = 0; f < F; ++f)
          tma_store_1d(a.x_out + fld[f] * a.pitch + (s - a.c0), out + fld[f] * WC + M4, static_cast<unsigned>(P * sizeof(double)));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      __syncwarp();
#pragma unroll
      for (int f = 0; f < F; ++f) {
        double* g = a.x_out + fld[f] * a.pitch + (s - a.c0);
        for (int j = lane; j < P; j += 32) g[j] = out[fld[f] * WC + M4 + j];
      }
      __syncwarp();
    }
  }
  if (tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (view.bad) *a.flag = 1;
}

}  // namespace odeintr_b200

extern "C" __global__ void __launch_bounds__(256, ODEINTR_WARP_MIN_BLOCKS)
odeintr_lattice_rk4_warp_h1(const odeintr_lattice_tiled_args a) { odeintr_b200::lattice_rk4_warp_body<1>(a); }

#if ODEINTR_WARP_CPT >= 2
extern "C" __global__ void __launch_bounds__(256, ODEINTR_WARP_MIN_BLOCKS)
odeintr_lattice_rk4_warp_h2(const odeintr_lattice_tiled_args a) { odeintr_b200::lattice_rk4_warp_body<2>(a); }
#endif

// cells per lane the kernels above were compiled with (the host sizes shared memory and ranges from it)
extern "C" __global__ void odeintr_lattice_warp_info(int* out) { out[0] = ODEINTR_WARP_CPT; }

#endif  // rk4 on a 1-D lattice
