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
// Two kernels share the work of one step:
//   odeintr_lattice_rk4_tiled           general tiles: ragged last tile, tiles at the periodic seam, tiles that
//                                       contain cells the user's loop does not write, slabs of a sharded run
//   odeintr_lattice_rk4_tiled_interior  all other tiles (all but two or three of them): register-blocked
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
  // tiles cover the cells this launch owns: [lo, hi) (the user loop's range on one GPU, the rank's slab otherwise)
  const long long lo = a.sharded ? a.lo : (start > a.lo ? start : a.lo);
  const long long hi = a.sharded ? a.hi : (bound < a.hi ? bound : a.hi);
  // the host hands the tiles [tile_first, tile_first + tile_skip) to the interior kernel below
  const long long tile = (int)blockIdx.x < a.tile_first ? (long long)blockIdx.x : (long long)blockIdx.x + a.tile_skip;
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
struct lattice_tile_rel_view {
  const double* p;
  long long n_total;  // cells per field
  int width;          // doubles between fields in shared memory
  int halo;
  int centre;         // global number of the cell being computed (generic path only)
  mutable int bad;    // a read reached beyond the halo or outside the state

  ODEINTR_DEV double at(long long f, long long d) const {
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
  // absolute index: no periodic image can be within reach of an interior cell, so the distance is taken as is
  ODEINTR_DEV double operator[](const long long g) const {
    const long long f = g >= 0 ? g / n_total : -1;
    return at(f, g - f * n_total - centre);
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(n_total); }
};

// Interior tiles.  Every thread owns CPT cells of the tile (tid + k * 256) for the whole step: the state at the start
// of the step and the running weighted sum of those cells never leave its registers; only the stage states, which the
// neighbours read, go through shared memory (two buffers, ping-pong).  The cells of the shrinking margins (3, 2, 1
// times halo on each side in stages 1, 2, 3) belong to the first 6 * halo threads, one each.
// Per cell and step: 1 LDG + 1 STG, 4 * (stencil reads) LDS + 4 STS -- against 20 LDS + 8 STS of the general body.
// Entry (field f, global cell c) of x_in / x_out sits at f * pitch + (c - c0): one GPU has pitch = L, c0 = 0; a rank
// of a sharded run its slab geometry (ghost cells at the negative offsets make every window contiguous).
ODEINTR_DEV void lattice_rk4_tiled_interior_body(const odeintr_lattice_tiled_args& a) {
  typedef odeintr::system_type_t<lattice_tile_rel_view> Sys;
  constexpr int F = ODEINTR_NFIELDS;
  constexpr int C = ODEINTR_TILE_CPT;
  constexpr int T = 256 * C;
  extern __shared__ double smem[];
  const int h = (int)a.halo;
  const int margin = 4 * h;
  const int W = T + 2 * margin;
  double* s0 = smem;          // x, then the stage-2 state
  double* s1 = smem + F * W;  // stage-1 state, then the stage-3 state
  const int tid = threadIdx.x;

  Sys sys;
  sys.load_params(a.pars, 1, 0);
  long long off[F];
  sys.field_offsets(off, 0);
  // cells per field, a compile-time constant here (the host checks it against its own count): the snippet's index
  // expressions can only fold against a constant lattice length
  constexpr long long L = static_cast<long long>(ODEINTR_SYS_SIZE) / ODEINTR_NFIELDS;
  const long long tile_lo = a.lo + ((long long)a.tile_first + blockIdx.x) * T;
  int fld[F];
#pragma unroll
  for (int f = 0; f < F; ++f) fld[f] = static_cast<int>(off[f] / L);

  // stage the window of x; the thread's own cells stay in registers as well
  double xs[C][F], acc[C][F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const double* g = a.x_in + fld[f] * a.pitch + (tile_lo - a.c0);
    double* s = s0 + fld[f] * W + margin;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const double v = __ldg(g + tid + k * 256);
      xs[k][f] = v;
      s[tid + k * 256] = v;
    }
    if (tid < 2 * margin) {
      const int j = tid < margin ? tid - margin : T + (tid - margin);
      s[j] = __ldg(g + j);
    }
  }
  __syncthreads();
  // the margin cell of this thread: je, at distance de + 1 from the tile, takes part in stage s while de < (4-s)*halo
  const bool has_edge = tid < 6 * h;
  const int de = tid < 3 * h ? tid : tid - 3 * h;
  const int je = tid < 3 * h ? -1 - tid : T + (tid - 3 * h);
  double xe[F];
#pragma unroll
  for (int f = 0; f < F; ++f) xe[f] = has_edge ? s0[fld[f] * W + margin + je] : 0.0;

  lattice_tile_rel_view view;
  view.n_total = L; view.width = W; view.halo = h; view.bad = 0;
  rel_index idx;
  idx.off = 0; idx.cells = static_cast<int>(L); idx.rel = true;

  const double half = 1.0 / 2.0, b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
  const double dt = a.dt;
#pragma unroll
  for (int stage = 1; stage <= 4; ++stage) {
    const double* src = (stage & 1) ? s0 : s1;
    double* dst = (stage & 1) ? s1 : s0;
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
      view.centre = idx.centre = static_cast<int>(tile_lo) + j;
      sys.cell(view, kk, idx, ts);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        if (stage < 4) dst[fld[f] * W + margin + j] = xs[k][f] + a_dt * kk[f];
        acc[k][f] = (stage == 1 ? xs[k][f] : acc[k][f]) + b_dt * kk[f];  // the running sum starts from x
        // new state of an owned cell: the only global store of the step
        if (stage == 4) a.x_out[fld[f] * a.pitch + (tile_lo - a.c0) + j] = acc[k][f];
      }
    }
    if (stage < 4) {
      if (has_edge && de < (4 - stage) * h) {
        double kk[F];
#pragma unroll
        for (int f = 0; f < F; ++f) kk[f] = 0.0;
        view.p = src + margin + je;
        view.centre = idx.centre = static_cast<int>(tile_lo) + je;
        sys.cell(view, kk, idx, ts);
#pragma unroll
        for (int f = 0; f < F; ++f) dst[fld[f] * W + margin + je] = xe[f] + a_dt * kk[f];
      }
      __syncthreads();
    }
  }
  if (view.bad) *a.flag = 1;
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

}  // namespace odeintr_b200

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

// what the host needs to know about the generated loop nest:
// out[0], out[1] = first cell and bound of the user's loop; out[2] = cells per field the interior kernel assumes;
// out[3] = 1 when every dxdt[...] target is `loop variable + constant` (the layout all lattice kernels write)
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
}

extern "C" __global__ void __launch_bounds__(256)
odeintr_lattice_rk4_tiled(const odeintr_lattice_tiled_args a) { odeintr_b200::lattice_rk4_tiled_body(a); }

extern "C" __global__ void __launch_bounds__(256)
odeintr_lattice_rk4_tiled_interior(const odeintr_lattice_tiled_args a) {
  odeintr_b200::lattice_rk4_tiled_interior_body(a);
}
