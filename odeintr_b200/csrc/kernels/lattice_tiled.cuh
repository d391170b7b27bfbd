// odeintr_b200/csrc/kernels/lattice_tiled.cuh
// K4t: a whole rk4 step of a large 1-D system in ONE kernel, temporally blocked in shared memory.
//
// The stage-per-launch schedule of lattice_rk.cuh moves 128 B per cell per step through HBM (every stage re-reads
// and re-writes the state-sized work arrays).  When the stencil radius `halo` of the user's loop is known (measured
// by odeintr_lattice_probe below, or declared through compile_sys_openmp(..., halo=)), a block can load its tile of
// T cells plus 4*halo cells on each side ONCE, run all four stages in shared memory on windows that shrink by
// `halo` per stage (the same redundant-edge scheme the multi-GPU path uses between ranks), and write only the new
// state: 16 B per cell per step, 8x less traffic, so the step becomes bound by the FP64 pipe instead of HBM.
// Every cell goes through exactly the operations of the staged schedule (Boost's order, no FMA): bit-identical.
//
// The radius is verified on EVERY access of every step: a read further than `halo` from the cell being computed
// raises a flag (and is served from a clamped position), and the host then repeats the integration with the staged
// kernels, so a snippet whose reach depends on the state or on time is still integrated correctly.
//
// ODEINTR_NO_SYMBOLIC (set by the host when the translation unit does not compile with the symbolic loop index, e.g.
// a snippet that passes the loop variable to a template expecting an int) leaves only the stage-per-launch kernels.
//
// Two kernels share the work of one step:
//   odeintr_lattice_rk4_tiled           general tiles: the cells at the periodic seam, at the ends of the user loop's
//                                       range (and cells the loop does not write), at the ends of a slab of a sharded
//                                       run, ragged ends
//   the interior cells (all but a few hundred)
//     radius 1 and 2:                   odeintr_lattice_rk4_warp_h1 / _h2 (lattice_warp.cuh): warp-autonomous, registers
//     wider stencils:                   odeintr_lattice_rk4_tiled_tma / _tiled_interior below: block tiles
#pragma once
#include "odeintr_b200/lattice_view.cuh"
#include "odeintr_b200/lattice_index.cuh"

#ifndef ODEINTR_TILE_CPT
#define ODEINTR_TILE_CPT 4  // cells per thread and stage: tile = 256 * 4 cells
#endif

namespace odeintr_b200 {

// General tiles: view of one stage's input held in shared memory, indexed with the snippet's GLOBAL indices.
struct lattice_tile_view {
  const double* s;        // [field][tile_cells + 2 * margin]
  long long n_total;      // cells per field (L < 2^31: the host refuses wider lattices for this kernel)
  int centre;             // global cell currently being computed (low 32 bits are all that matters below)
  int cells;              // (int) n_total
  int j;                  // window index of that cell (0 = first tile cell)
  int width;              // tile_cells + 2 * margin
  int margin;             // 4 * halo
  int halo;
  int* flag;              // global error word

  // Only the DISTANCE of the requested entry from the cell being computed matters, and it fits 32 bits, so the
  // high halves of the snippet's 64-bit index expressions are dead code for the compiler.
  ODEINTR_DEV double operator[](const long long g) const {
#if defined(ODEINTR_NFIELDS) && ODEINTR_NFIELDS > 1
    int f = 0;
#pragma unroll
    for (int q = 1; q < ODEINTR_NFIELDS; ++q)
      if (g >= q * n_total) f = q;
    int r = static_cast<int>(g - f * n_total) - centre;
#else
    const int f = 0;
    int r = static_cast<int>(g) - centre;
#endif
    // a neighbour across the periodic seam is L - r cells away in index space
    if (r > halo) r -= cells; else if (r < -halo) r += cells;
    if (g < 0 || g >= ODEINTR_NFIELDS * n_total || r > halo || r < -halo) { *flag = 1; return 0.0; }
    return s[f * width + j + r + margin];
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(n_total); }
};

#if !defined(ODEINTR_LATTICE_2D) && !defined(ODEINTR_NO_SYMBOLIC)  // M x M lattices: lattice_tiled2d.cuh

ODEINTR_DEV void lattice_rk4_tiled_body(const odeintr_lattice_tiled_args& a) {
  typedef odeintr::system_type_t<lattice_tile_view> Sys;
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int T = 256 * ODEINTR_TILE_CPT;
  extern __shared__ double smem[];
  const int margin = 4 * (int)a.halo;
  const int W = T + 2 * margin;
  double* sx = smem;               // state at the start of the step
  double* sa = sx + F * W;         // stage states, ping
  double* sb = sa + F * W;         // stage states, pong
  double* sacc = sb + F * W;       // running weighted sum

  Sys sys;
  sys.load_params(a.pars, 1, 0);
  long long off[F];
  sys.field_offsets(off, 0);
  const long long start = sys.cell_start(), bound = sys.cell_bound();
  const long long L = a.n_total;
  // tiles cover the two ranges of cells this launch owns, [lo, hi) and [lo2, hi2): what the host did not hand to the
  // interior / warp kernels (the ends of the user loop's range on one GPU, of the rank's slab otherwise)
  const long long tiles_a = a.hi > a.lo ? (a.hi - a.lo + T - 1) / T : 0;
  const bool second = static_cast<long long>(blockIdx.x) >= tiles_a;
  const long long lo = second ? a.lo2 : a.lo, hi = second ? a.hi2 : a.hi;
  const long long tile = second ? static_cast<long long>(blockIdx.x) - tiles_a : static_cast<long long>(blockIdx.x);
  const long long tile_lo = lo + tile * T;
  if (tile_lo >= hi) return;
  const long long tile_hi = tile_lo + T < hi ? tile_lo + T : hi;
  const int tn = (int)(tile_hi - tile_lo);

  int fld[F];
  long long out_base[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    fld[f] = static_cast<int>(off[f] / L);  // field number of the f-th written entry
    out_base[f] = a.sharded ? (off[f] / L) * a.pitch - a.c0 : off[f];
  }

  // load x on [tile_lo - margin, tile_lo + T + margin) of every field, periodic in the cell index; cells outside the
  // user loop keep their value in every stage, so all four arrays start as copies of x
  lattice_view_t<LATTICE_WRAP> gw;
  gw.p = a.x_in; gw.n_total = L; gw.n_local = a.n; gw.c0 = a.c0; gw.ghost = a.ghost; gw.pitch = a.pitch;
  const int wn = tn + 2 * margin;  // a ragged tile stages only what it needs (a slab's buffer ends there)
  for (int f = 0; f < F; ++f) {
    for (int j = threadIdx.x; j < wn; j += blockDim.x) {
      long long c = tile_lo - margin + j;
      if (c < 0) c += L; else if (c >= L) c -= L;
      const double v = a.sharded ? gw[fld[f] * L + c] : __ldg(a.x_in + fld[f] * L + c);
      sx[fld[f] * W + j] = v;
      sa[fld[f] * W + j] = v;
      sb[fld[f] * W + j] = v;
      sacc[fld[f] * W + j] = v;
    }
  }
  __syncthreads();

  lattice_tile_view view;
  view.n_total = L; view.cells = static_cast<int>(L); view.width = W; view.margin = margin; view.halo = (int)a.halo;
  view.flag = a.flag;

  const double half = 1.0 / 2.0, b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
  const double dt = a.dt;
#pragma unroll 1
  for (int stage = 1; stage <= 4; ++stage) {
    // stage s works on the window [-(4-s)*halo, tn + (4-s)*halo) around the tile
    const int m = (4 - stage) * (int)a.halo;
    const double* src = stage == 1 ? sx : (stage == 2 ? sa : (stage == 3 ? sb : sa));
    double* dst = stage == 1 ? sa : (stage == 2 ? sb : sa);  // stage 3 overwrites sa, which stage 2 has finished reading
    const double a_dt = stage == 1 ? half * dt : (stage == 2 ? half * dt : 1.0 * dt);
    const double b_dt = stage == 1 ? b1 * dt : (stage == 2 ? b2 * dt : (stage == 3 ? b2 * dt : b1 * dt));
    const double ts = stage == 1 ? a.t : (stage == 4 ? a.t + 1.0 * dt : a.t + half * dt);
    view.s = src;
    for (int j = threadIdx.x - m; j < tn + m; j += blockDim.x) {
      long long cg = tile_lo + j;  // global cell, possibly a periodic image
      if (cg < 0) cg += L; else if (cg >= L) cg -= L;
      if (cg < start || cg >= bound) continue;  // not a cell of the user's loop: keeps its value
      double k[F];
#pragma unroll
      for (int f = 0; f < F; ++f) k[f] = 0.0;
      view.centre = static_cast<int>(cg);
      view.j = j;
      sys.cell(view, k, cg, ts);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const int q = fld[f] * W + j + margin;
        if (stage == 1) {
          dst[q] = sx[q] + a_dt * k[f];
          sacc[q] = sx[q] + b_dt * k[f];  // the running sum starts from x
        } else if (stage < 4) {
          dst[q] = sx[q] + a_dt * k[f];
          sacc[q] = sacc[q] + b_dt * k[f];
        } else {
          // new state of an owned cell: the only global store of the step
          a.x_out[out_base[f] + (tile_lo + j)] = sacc[q] + b_dt * k[f];
        }
      }
    }
    __syncthreads();
  }
}

// View of one stage's input for a cell of an INTERIOR tile (full tile, whole window inside the user loop's range and
// away from the periodic seam).  Indexed with a rel_index, the address is `p + constant`: p points at the entry
// (field 0, cell being computed) of the stage's input, the displacement folds at compile time -- and so does the
// reach check against the (block-uniform) halo.
template <bool RING>
struct lattice_rel_view {
  const double* p;
  long long n_total;  // cells per field
  int width;          // doubles between fields in shared memory
  int halo;
  int centre;         // global number of the cell being computed (generic path only)
  mutable int bad;    // a read reached beyond the halo or outside the state
  const double* own;  // the stage's input at the cell being computed, per field: the thread wrote it itself

  ODEINTR_DEV double at(long long f, long long d) const {
    // the cell's own entries come from registers (a compile-time test for ordinary stencils)
    if (d == 0 && f >= 0 && f < ODEINTR_NFIELDS) return own[f];
    // out-of-reach reads are served from a clamped position (never outside the staged window) and reported once
    // per thread at the end of the kernel
    if (f < 0) { bad = 1; f = 0; }
    if (f >= ODEINTR_NFIELDS) { bad = 1; f = ODEINTR_NFIELDS - 1; }
    if (d > halo) { bad = 1; d = halo; }
    if (d < -halo) { bad = 1; d = -halo; }
    return p[f * width + d];
  }
  ODEINTR_DEV double operator[](const rel_index& r) const {
    if (!r.rel) return (*this)[r.off];
    // off = f * L + d with |d| small: nearest multiple of L
    const long long L = n_total;
    const long long f = r.off >= 0 ? (r.off + L / 2) / L : -((-r.off + L / 2) / L);
    return at(f, r.off - f * L);
  }
  // absolute index.  Tile view: no periodic image can be within reach of an interior cell, so the distance is taken
  // as is.  RING (the whole lattice is staged, ghost cells hold the periodic images): nearest image.
  ODEINTR_DEV double operator[](const long long g) const {
    const long long f = g >= 0 ? g / n_total : -1;
    long long d = g - f * n_total - centre;
    if (RING) {
      if (d > n_total / 2) d -= n_total;
      else if (d < -(n_total / 2)) d += n_total;
    }
    return at(f, d);
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(n_total); }
};
typedef lattice_rel_view<false> lattice_tile_rel_view;

// Interior tiles.  Every thread owns CPT cells of the tile (tid + k * 256) for the whole step: the state at the start
// of the step, the running weighted sum and the current stage state of those cells never leave its registers; the
// stage states also go to shared memory, where the neighbours read them.  The cells of the shrinking margins (3, 2, 1
// times halo on each side in stages 1, 2, 3) belong to the first 6 * halo threads, one each.
// Per cell and step: 8 bytes in + 8 bytes out of HBM, 4 * (stencil reads other than the cell itself) LDS + 3 STS --
// against 20 LDS + 8 STS of the general body.
// Entry (field f, global cell c) of x_in / x_out sits at f * pitch + (c - c0): one GPU has pitch = L, c0 = 0; a rank
// of a sharded run its slab geometry (ghost cells at the negative offsets make every window contiguous).
struct lattice_tile_ctx {
  odeintr::system_type_t<lattice_tile_rel_view> sys;
  int fld[ODEINTR_NFIELDS];  // field number of the f-th written entry
  int h, margin, W;
};

// One tile whose window of x is already staged in `in` ([field][W]).  Stage states: 1 -> sa, 2 -> in, 3 -> sb
// (sb may be sa), 4 -> x_out.  Ends without a barrier: the last stage only reads sb.
template <bool TMA_INPUT>
ODEINTR_DEV void lattice_rk4_tile(const odeintr_lattice_tiled_args& a, lattice_tile_ctx& cx, double* in, double* sa,
                                  double* sb, const long long tile_lo, int& bad) {
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int C = ODEINTR_TILE_CPT;
  constexpr int T = 256 * C;
  constexpr long long L = static_cast<long long>(ODEINTR_SYS_SIZE) / ODEINTR_NFIELDS;
  const int tid = threadIdx.x, h = cx.h, margin = cx.margin, W = cx.W;

  // the thread's own cells, and its margin cell je: at distance de + 1 from the tile, it takes part in stage s while
  // de < (4 - s) * halo
  double xs[C][F], acc[C][F], xe[F];
  const bool has_edge = tid < 6 * h;
  const int de = tid < 3 * h ? tid : tid - 3 * h;
  const int je = tid < 3 * h ? -1 - tid : T + (tid - 3 * h);
#pragma unroll
  for (int f = 0; f < F; ++f) {
#pragma unroll
    for (int k = 0; k < C; ++k) xs[k][f] = in[cx.fld[f] * W + margin + tid + k * 256];
    xe[f] = has_edge ? in[cx.fld[f] * W + margin + je] : 0.0;
  }

  lattice_tile_rel_view view;
  view.n_total = L; view.width = W; view.halo = h; view.bad = 0;
  rel_index idx;
  idx.off = 0; idx.cells = static_cast<int>(L); idx.rel = true;

  // the stage's input at the thread's own cells, by field number (what the thread stored in the previous stage)
  double own[C][F], own_e[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
#pragma unroll
    for (int k = 0; k < C; ++k) own[k][cx.fld[f]] = xs[k][f];
    own_e[cx.fld[f]] = xe[f];
  }

  const double half = 1.0 / 2.0, b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
  const double dt = a.dt;
#pragma unroll
  for (int stage = 1; stage <= 4; ++stage) {
    const double* src = stage == 1 ? in : (stage == 2 ? sa : (stage == 3 ? in : sb));
    double* dst = stage == 1 ? sa : (stage == 2 ? in : sb);
    const double a_dt = stage == 1 ? half * dt : (stage == 2 ? half * dt : 1.0 * dt);
    const double b_dt = stage == 1 ? b1 * dt : (stage == 2 ? b2 * dt : (stage == 3 ? b2 * dt : b1 * dt));
    const double ts = stage == 1 ? a.t : (stage == 4 ? a.t + 1.0 * dt : a.t + half * dt);
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const int j = tid + k * 256;
      double kk[F];
#pragma unroll
      for (int f = 0; f < F; ++f) kk[f] = 0.0;
      view.p = src + margin + j;
      view.own = own[k];
      view.centre = idx.centre = static_cast<int>(tile_lo) + j;
      cx.sys.cell(view, kk, idx, ts);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        if (stage < 4) {
          const double v = xs[k][f] + a_dt * kk[f];
          dst[cx.fld[f] * W + margin + j] = v;
          own[k][cx.fld[f]] = v;
        }
        acc[k][f] = (stage == 1 ? xs[k][f] : acc[k][f]) + b_dt * kk[f];  // the running sum starts from x
        // new state of an owned cell: the only global store of the step
        if (stage == 4) a.x_out[cx.fld[f] * a.pitch + (tile_lo - a.c0) + j] = acc[k][f];
      }
    }
    if (stage < 4) {
      if (has_edge && de < (4 - stage) * h) {
        double kk[F];
#pragma unroll
        for (int f = 0; f < F; ++f) kk[f] = 0.0;
        view.p = src + margin + je;
        view.own = own_e;
        view.centre = idx.centre = static_cast<int>(tile_lo) + je;
        cx.sys.cell(view, kk, idx, ts);
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const double v = xe[f] + a_dt * kk[f];
          dst[cx.fld[f] * W + margin + je] = v;
          own_e[cx.fld[f]] = v;
        }
      }
      // stage 2 stores into the buffer the TMA engine refills later: make them visible to the async proxy
      if (TMA_INPUT && stage == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
  }
  bad |= view.bad;
}

ODEINTR_DEV void lattice_tile_ctx_init(const odeintr_lattice_tiled_args& a, lattice_tile_ctx& cx) {
  constexpr int F = ODEINTR_NFIELDS;
  // cells per field, a compile-time constant here (the host checks it against its own count): the snippet's index
  // expressions can only fold against a constant lattice length
  constexpr long long L = static_cast<long long>(ODEINTR_SYS_SIZE) / ODEINTR_NFIELDS;
  cx.sys.load_params(a.pars, 1, 0);
  long long off[F];
  cx.sys.field_offsets(off, 0);
#pragma unroll
  for (int f = 0; f < F; ++f) cx.fld[f] = static_cast<int>(off[f] / L);
  cx.h = (int)a.halo;
  cx.margin = 4 * cx.h;
  cx.W = 256 * ODEINTR_TILE_CPT + 2 * cx.margin;
}

// one block per tile, window staged with ordinary loads (any alignment)
ODEINTR_DEV void lattice_rk4_tiled_interior_body(const odeintr_lattice_tiled_args& a) {
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int T = 256 * ODEINTR_TILE_CPT;
  extern __shared__ double smem[];
  lattice_tile_ctx cx;
  lattice_tile_ctx_init(a, cx);
  const int tid = threadIdx.x, margin = cx.margin, W = cx.W;
  double* s0 = smem;          // x, then the stage-2 state
  double* s1 = smem + F * W;  // stage-1 state, then the stage-3 state
  const long long tile_lo = a.lo + ((long long)a.tile_first + blockIdx.x) * T;
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const double* g = a.x_in + cx.fld[f] * a.pitch + (tile_lo - a.c0);
    double* s = s0 + cx.fld[f] * W + margin;
    double v[ODEINTR_TILE_CPT];  // all loads in flight before the first store
#pragma unroll
    for (int k = 0; k < ODEINTR_TILE_CPT; ++k) v[k] = __ldg(g + tid + k * 256);
    if (tid < 2 * margin) {
      const int j = tid < margin ? tid - margin : T + (tid - margin);
      s[j] = __ldg(g + j);
    }
#pragma unroll
    for (int k = 0; k < ODEINTR_TILE_CPT; ++k) s[tid + k * 256] = v[k];
  }
  __syncthreads();
  int bad = 0;
  lattice_rk4_tile<false>(a, cx, s0, s1, s1, tile_lo, bad);
  if (bad) *a.flag = 1;
}

// ---- TMA-staged, persistent variant ----------------------------------------------------------------------------
// One block per SM slot walks over the interior tiles b, b + gridDim.x, ...  The window of the NEXT tile is fetched
// by the TMA engine (cp.async.bulk global -> shared, one bulk copy per field, completion counted in bytes on an
// mbarrier) into the second input buffer while the block computes the current tile, so no warp ever waits for HBM.
// Needs 16-byte aligned windows (the host checks: even window start, even pitch); otherwise the kernel above runs.
ODEINTR_DEV unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

ODEINTR_DEV void mbar_init(unsigned long long* bar, const unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
ODEINTR_DEV void mbar_expect_tx(unsigned long long* bar, const unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
ODEINTR_DEV void mbar_wait(unsigned long long* bar, const unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
}
ODEINTR_DEV void tma_load_1d(void* dst, const void* src, const unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

ODEINTR_DEV void lattice_rk4_tiled_tma_body(const odeintr_lattice_tiled_args& a) {
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int T = 256 * ODEINTR_TILE_CPT;
  extern __shared__ __align__(128) double smem_tma[];
  double* smem = smem_tma;
  lattice_tile_ctx cx;
  lattice_tile_ctx_init(a, cx);
  const int tid = threadIdx.x, margin = cx.margin, W = cx.W;
  // input windows of the current and the next tile (then stage-2 states) at smem + b * F * W, b = 0, 1
  double* s1 = smem + 2 * F * W;         // stage-1 states
  double* s2 = smem + 3 * F * W;         // stage-3 states: a buffer of their own, so that a warp starting the next
                                         // tile's stage 1 cannot overwrite what a slower warp still reads in stage 4
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 4 * F * W);
  const long long n_tiles = a.tile_skip;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned bytes = static_cast<unsigned>(W * sizeof(double));
  int bad = 0;
  int it = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = it & 1;
    if (tid == 0) {
      // one thread asks the TMA engine for the window of a tile: -> input buffer nb, completion on bar[nb].
      // First pass: this tile's own window as well.  Every thread has passed the barrier that ends stage 3 of the
      // previous tile, the last use of the other buffer; the proxy fence orders those ordinary shared-memory
      // accesses before the engine's writes.
      for (int ahead = (it == 0 ? 0 : 1); ahead <= 1; ++ahead) {
        const long long tl = tile + ahead * (long long)gridDim.x;
        if (tl >= n_tiles) break;
        const int nb = b ^ ahead;
        const long long lo = a.lo + (a.tile_first + tl) * T - margin;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar + nb, bytes * F);
#pragma unroll
        for (int f = 0; f < F; ++f)
          tma_load_1d(smem + (nb * F + cx.fld[f]) * W, a.x_in + cx.fld[f] * a.pitch + (lo - a.c0), bytes, bar + nb);
      }
    }
    mbar_wait(bar + b, (it >> 1) & 1);
    lattice_rk4_tile<true>(a, cx, smem + b * F * W, s1, s2, a.lo + (a.tile_first + tile) * T, bad);
  }
  if (bad) *a.flag = 1;
}

// Reach probe: evaluates the loop body of every cell against a view that reads nothing and records how far from the
// cell each x[...] lies (periodic distance within its field).  The maximum is the stencil radius the tiled kernels
// need; an index outside the state reports a radius no tile can hold.
struct lattice_probe_view {
  long long n_total;
  int centre;
  mutable int reach;
  ODEINTR_DEV void note(const long long f, long long d) const {
    if (d < 0) d = -d;
    if (d > n_total - d) d = n_total - d;
    if (f < 0 || f >= ODEINTR_NFIELDS) d = 1 << 30;
    if (d > reach) reach = static_cast<int>(d < (1 << 30) ? d : (1 << 30));
  }
  ODEINTR_DEV double operator[](const rel_index& r) const {
    if (!r.rel) return (*this)[r.off];
    const long long L = n_total;
    const long long f = r.off >= 0 ? (r.off + L / 2) / L : -((-r.off + L / 2) / L);
    note(f, r.off - f * L);
    return 1.0;
  }
  ODEINTR_DEV double operator[](const long long g) const {
    const long long f = g >= 0 ? g / n_total : -1;
    note(f, g - f * n_total - centre);
    return 1.0;
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(n_total); }
};

#endif  // !ODEINTR_LATTICE_2D

}  // namespace odeintr_b200

#if !defined(ODEINTR_LATTICE_2D) && !defined(ODEINTR_NO_SYMBOLIC)
// out[0] = largest distance (in cells) between a cell and an entry its loop body reads
extern "C" __global__ void __launch_bounds__(256)
odeintr_lattice_probe(const double* pars, const double t, int* out) {
  typedef odeintr::system_type_t<odeintr_b200::lattice_probe_view> Sys;
  constexpr long long L = static_cast<long long>(ODEINTR_SYS_SIZE) / ODEINTR_NFIELDS;
  Sys sys;
  sys.load_params(pars, 1, 0);
  odeintr_b200::lattice_probe_view view;
  view.n_total = L; view.reach = 0;
  odeintr_b200::rel_index idx;
  idx.off = 0; idx.cells = static_cast<int>(L); idx.rel = true;
  const long long start = sys.cell_start() > 0 ? sys.cell_start() : 0;
  const long long bound = sys.cell_bound() < L ? sys.cell_bound() : L;
  for (long long c = start + blockIdx.x * (long long)blockDim.x + threadIdx.x; c < bound;
       c += (long long)gridDim.x * blockDim.x) {
    double k[ODEINTR_NFIELDS];
    view.centre = idx.centre = static_cast<int>(c);
    sys.cell(view, k, idx, t);
  }
  if (view.reach) atomicMax(out, view.reach);
}

#endif  // !ODEINTR_LATTICE_2D

// what the host needs to know about the generated loop nest:
// out[0], out[1] = first cell and bound of the user's loop; out[2] = cells per field the interior kernel assumes;
// out[3] = 1 when every dxdt[...] target is `loop variable + constant` (the layout all lattice kernels write);
// out[4] = M when the snippet addresses an M x M lattice (one field), else 0
extern "C" __global__ void odeintr_lattice_info(long long* out) {
  odeintr::system_type_t<odeintr_b200::lattice_tile_view> sys;
  out[0] = sys.cell_start();
  out[1] = sys.cell_bound();
  out[2] = static_cast<long long>(ODEINTR_SYS_SIZE) / ODEINTR_NFIELDS;
  long long o0[ODEINTR_NFIELDS], o1[ODEINTR_NFIELDS], o7[ODEINTR_NFIELDS];
  sys.field_offsets(o0, 0);
  sys.field_offsets(o1, 1);
  sys.field_offsets(o7, 7);
  long long unit = 1;
  for (int f = 0; f < ODEINTR_NFIELDS; ++f)
    if (o1[f] - o0[f] != 1 || o7[f] - o0[f] != 7) unit = 0;
  out[3] = unit;
#if defined(ODEINTR_LATTICE_2D) && ODEINTR_NFIELDS == 1
  out[4] = odeintr_b200::isqrt(ODEINTR_SYS_SIZE);
#else
  out[4] = 0;
#endif
}

#if !defined(ODEINTR_LATTICE_2D) && !defined(ODEINTR_NO_SYMBOLIC) && ODEINTR_LATTICE_STAGES == 4
extern "C" __global__ void __launch_bounds__(256)
odeintr_lattice_rk4_tiled(const odeintr_lattice_tiled_args a) { odeintr_b200::lattice_rk4_tiled_body(a); }

#ifndef ODEINTR_TILE_MIN_BLOCKS  // resident blocks per SM the interior kernel is compiled for (register cap)
#define ODEINTR_TILE_MIN_BLOCKS (ODEINTR_NFIELDS > 1 ? 3 : 4)
#endif

extern "C" __global__ void __launch_bounds__(256, ODEINTR_TILE_MIN_BLOCKS)
odeintr_lattice_rk4_tiled_interior(const odeintr_lattice_tiled_args a) {
  odeintr_b200::lattice_rk4_tiled_interior_body(a);
}

#ifndef ODEINTR_TMA_MIN_BLOCKS  // the persistent kernel hides latency by prefetching, not by occupancy
#define ODEINTR_TMA_MIN_BLOCKS (ODEINTR_NFIELDS > 1 ? 2 : 4)
#endif

extern "C" __global__ void __launch_bounds__(256, ODEINTR_TMA_MIN_BLOCKS)
odeintr_lattice_rk4_tiled_tma(const odeintr_lattice_tiled_args a) { odeintr_b200::lattice_rk4_tiled_tma_body(a); }
#endif  // rk4 on a 1-D lattice
