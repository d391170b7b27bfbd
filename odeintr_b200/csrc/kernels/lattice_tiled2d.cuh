// odeintr_b200/csrc/kernels/lattice_tiled2d.cuh
// K4t2: a whole rk4 step of a periodic M x M lattice in ONE kernel pass -- the temporally blocked step of
// lattice_tiled.cuh for the reference's own large-system example (compile_sys_openmp with laplace2D / laplace4,
// /root/reference R/odeintr.R:393-413, inst/include/utils.h:6-31).
//
// A block owns a tile of TH x TW cells.  It stages the tile plus 4 * halo cells on every side (rows and columns
// wrapped periodically while loading, so the window in shared memory is plain and every tile is treated alike), runs
// the four stages on windows that shrink by `halo` per stage and writes the new state of its tile: 8 bytes read
// (x 1.4 for the margins) and 8 written per cell and step instead of the 128 of the stage-per-launch schedule.
// Every thread owns TH/4 consecutive cells of one tile column: their x, running sum and stage state stay in registers
// (so the rows above / below are register reads too).  The
// margin ("frame") cells of stages 1-3 are dealt round-robin, their x is kept in registers too; two shared-memory
// windows (x / stage 2, stage 1 / stage 3) carry what the neighbours read.
// The snippet's coordinates are symbolic (lattice_index.cuh): `laplace2D(x, i / M, i % M)` compiles to four
// shared-memory loads at fixed (row, column) offsets from the thread's cell.
// Bit-identical to the staged schedule; every access is checked against the radius (see lattice_tiled.cuh).
#pragma once
#include "odeintr_b200/lattice_view.cuh"
#include "odeintr_b200/lattice_index.cuh"

#if defined(ODEINTR_LATTICE_2D) && ODEINTR_NFIELDS == 1 && !defined(ODEINTR_NO_SYMBOLIC)

#ifndef ODEINTR_TILE2_ROWS
#define ODEINTR_TILE2_ROWS 32
#endif
#define ODEINTR_TILE2_COLS 64

namespace odeintr_b200 {

// nearest periodic image of a coordinate difference
ODEINTR_DEV long long min_image(long long d, const long long m) {
  if (d > m / 2) d -= m;
  else if (d < -(m / 2)) d += m;
  return d;
}

// One stage's input around the cell being computed, in the shared-memory window (row pitch `pitch`).
struct lattice_tile2_view {
  const double* p;  // the cell being computed
  double own;       // its value
  const double* mine;  // the stage's input at the thread's own cells (consecutive rows of one column), or null
  int k, n_mine;       // which of them is being computed
  int pitch, halo;
  int row, col, side;
  mutable int bad;

  ODEINTR_DEV double at2(long long dr, long long dc) const {
    if (dr == 0 && dc == 0) return own;
    // the rows above / below in the thread's own column strip: registers (compile-time test after unrolling)
    if (dc == 0 && mine && k + dr >= 0 && k + dr < n_mine) return mine[k + dr];
    if (dr > halo) { bad = 1; dr = halo; }
    if (dr < -halo) { bad = 1; dr = -halo; }
    if (dc > halo) { bad = 1; dc = halo; }
    if (dc < -halo) { bad = 1; dc = -halo; }
    return p[dr * pitch + dc];
  }
  ODEINTR_DEV double operator[](const rel_cell2& c) const {
    if (c.rel && c.has_row && c.has_col) return at2(c.dr, c.dc);
    return (*this)[c.abs];
  }
  ODEINTR_DEV double operator[](const rel_index2&) const { return own; }
  ODEINTR_DEV double operator[](const long long g) const {
    if (g < 0 || g >= static_cast<long long>(side) * side) { bad = 1; return own; }
    const long long r = g / side;
    return at2(min_image(r - row, side), min_image(g - r * side - col, side));
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(side) * side; }
};

// reach probe for M x M lattices: the largest row or column distance the loop body reads
struct lattice_probe2_view {
  int row, col, side;
  mutable int reach;
  ODEINTR_DEV void note(long long dr, long long dc) const {
    if (dr < 0) dr = -dr;
    if (dc < 0) dc = -dc;
    const long long d = dr > dc ? dr : dc;
    if (d > reach) reach = static_cast<int>(d < (1 << 30) ? d : (1 << 30));
  }
  ODEINTR_DEV double operator[](const rel_cell2& c) const {
    if (c.rel && c.has_row && c.has_col) { note(c.dr, c.dc); return 1.0; }
    return (*this)[c.abs];
  }
  ODEINTR_DEV double operator[](const rel_index2&) const { return 1.0; }
  ODEINTR_DEV double operator[](const long long g) const {
    if (g < 0 || g >= static_cast<long long>(side) * side) { note(1 << 30, 0); return 1.0; }
    const long long r = g / side;
    note(min_image(r - row, side), min_image(g - r * side - col, side));
    return 1.0;
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(side) * side; }
};

}  // namespace odeintr_b200

extern "C" __global__ void __launch_bounds__(256)
odeintr_lattice_probe(const double* pars, const double t, int* out) {
  typedef odeintr::system_type_t<odeintr_b200::lattice_probe2_view> Sys;
  constexpr int MM = odeintr_b200::isqrt(ODEINTR_SYS_SIZE);
  Sys sys;
  sys.load_params(pars, 1, 0);
  odeintr_b200::lattice_probe2_view view;
  view.side = MM; view.reach = 0;
  odeintr_b200::rel_index2 idx;
  idx.side = MM;
  const long long start = sys.cell_start() > 0 ? sys.cell_start() : 0;
  const long long bound = sys.cell_bound() < (long long)MM * MM ? sys.cell_bound() : (long long)MM * MM;
  for (long long c = start + blockIdx.x * (long long)blockDim.x + threadIdx.x; c < bound;
       c += (long long)gridDim.x * blockDim.x) {
    double k[1];
    idx.centre = static_cast<int>(c);
    view.row = idx.row = static_cast<int>(c / MM);
    view.col = idx.col = static_cast<int>(c - (long long)idx.row * MM);
    sys.cell(view, k, idx, t);
  }
  if (view.reach) atomicMax(out, view.reach);
}

#ifndef ODEINTR_TILE2_MIN_BLOCKS
#define ODEINTR_TILE2_MIN_BLOCKS 3
#endif

namespace odeintr_b200 {

// H = stencil radius, a compile-time constant: window pitch, frame geometry and every shared-memory offset fold
template <int H>
ODEINTR_DEV void lattice_rk4_tiled2d_body(const odeintr_lattice_tiled_args& a) {
  typedef odeintr::system_type_t<lattice_tile2_view> Sys;
  constexpr int MM = isqrt(ODEINTR_SYS_SIZE);
  constexpr int TH = ODEINTR_TILE2_ROWS, TW = ODEINTR_TILE2_COLS;
  constexpr int C = TH / 4;  // cells per thread: rows ty * C ... ty * C + C - 1 of column tx
  constexpr int mg = 4 * H;
  constexpr int PW = TW + 2 * mg, PH = TH + 2 * mg;
  extern __shared__ double smem[];
  const int tid = threadIdx.x, tx = tid & (TW - 1), ty = tid / TW;
  double* const P = smem;             // x on the window, then the stage-2 states
  double* const A = P + PH * PW;      // stage-1 states, then stage-3 states

  Sys sys;
  sys.load_params(a.pars, 1, 0);
  const long long start = sys.cell_start(), bound = sys.cell_bound();
  const bool all_live = start <= 0 && bound >= (long long)MM * MM;
  constexpr int tiles_x = (MM + TW - 1) / TW;
  const int r0 = (int)(blockIdx.x / tiles_x) * TH, c0 = (int)(blockIdx.x % tiles_x) * TW;

  auto wrap = [&](int v) {  // usually one step: coordinates reach at most one tile beyond the lattice
    while (v < 0) v += MM;
    while (v >= MM) v -= MM;
    return v;
  };
  // stage the window, rows and columns wrapped periodically; all loads in flight before the first store
  {
    constexpr int NL = (PH * PW + 255) / 256;
    double v[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const int q = tid + k * 256;
      const int y = q / PW, x = q - y * PW;
      v[k] = q < PH * PW ? __ldg(a.x_in + (long long)wrap(r0 - mg + y) * MM + wrap(c0 - mg + x)) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const int q = tid + k * 256;
      if (q < PH * PW) {
        P[q] = v[k];
        if (!all_live) A[q] = v[k];  // cells the loop does not write keep their value in every stage
      }
    }
  }
  __syncthreads();

  lattice_tile2_view view;
  view.pitch = PW; view.halo = H; view.side = MM; view.bad = 0;
  rel_index2 idx;
  idx.side = MM;
  auto at_cell = [&](const int y, const int x) {  // window-relative cell (y, x), tile origin at (0, 0)
    view.row = idx.row = wrap(r0 + y);
    view.col = idx.col = wrap(c0 + x);
    idx.centre = idx.row * MM + idx.col;
  };

  // the thread's own cells of the tile ...
  double xs[C], acc[C], own[2][C];  // own[s & 1]: the input of stage s at the thread's cells
  bool live[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    const int y = ty * C + k;
    xs[k] = own[1][k] = P[(mg + y) * PW + mg + tx];
    acc[k] = xs[k];
    const long long flat = (long long)wrap(r0 + y) * MM + wrap(c0 + tx);
    live[k] = all_live || (flat >= start && flat < bound);
  }
  // ... and its cells of the frame of width 3 H around the tile (stage 1 computes all of it, stages 2 and 3 the
  // inner 2 H and H): top and bottom bands of full width first, then the left and right bands
  constexpr int FM = 3 * H, FW = TW + 2 * FM;
  constexpr int N_BAND = FM * FW, N_SIDE = TH * FM, N_FRAME = 2 * N_BAND + 2 * N_SIDE;
  constexpr int NF = (N_FRAME + 255) / 256;
  double fx[NF];
  int fw[NF], fd[NF];  // window offset, distance from the tile (0 = not a frame cell of this thread)
#pragma unroll
  for (int e = 0; e < NF; ++e) {
    const int q = tid + e * 256;
    int y, x;
    if (q < 2 * N_BAND) {
      const int b = q < N_BAND ? q : q - N_BAND;
      const int yy = b / FW;
      y = q < N_BAND ? yy - FM : TH + yy;
      x = b - yy * FW - FM;
    } else {
      const int sq = q - 2 * N_BAND;
      const int b = sq < N_SIDE ? sq : sq - N_SIDE;
      y = b / FM;
      const int xx = b - y * FM;
      x = sq < N_SIDE ? xx - FM : TW + xx;
    }
    const int dy = y < 0 ? -y : (y >= TH ? y - (TH - 1) : 0), dx = x < 0 ? -x : (x >= TW ? x - (TW - 1) : 0);
    fw[e] = (mg + y) * PW + mg + x;
    fd[e] = q < N_FRAME ? (dy > dx ? dy : dx) : 0;
    fx[e] = 0.0;
    if (fd[e]) {
      fx[e] = P[fw[e]];
      if (!all_live) {
        at_cell(y, x);
        if (!(idx.centre >= start && idx.centre < bound)) fd[e] = 0;  // keeps its value
      }
    }
  }

  const double half = 1.0 / 2.0, b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
  const double dt = a.dt;
#pragma unroll
  for (int stage = 1; stage <= 4; ++stage) {
    const double* src = (stage & 1) ? P : A;
    double* dst = (stage & 1) ? A : P;
    const double a_dt = stage == 1 ? half * dt : (stage == 2 ? half * dt : 1.0 * dt);
    const double b_dt = stage == 1 ? b1 * dt : (stage == 2 ? b2 * dt : (stage == 3 ? b2 * dt : b1 * dt));
    const double ts = stage == 1 ? a.t : (stage == 4 ? a.t + 1.0 * dt : a.t + half * dt);
#pragma unroll
    for (int k = 0; k < C; ++k) {
      if (!live[k]) {
        own[(stage + 1) & 1][k] = xs[k];  // keeps its value in every stage
        continue;
      }
      const int y = ty * C + k;
      const int q = (mg + y) * PW + mg + tx;
      double kk[1] = {0.0};
      view.p = src + q;
      view.own = own[stage & 1][k];
      view.mine = own[stage & 1]; view.k = k; view.n_mine = C;
      at_cell(y, tx);
      sys.cell(view, kk, idx, ts);
      acc[k] = (stage == 1 ? xs[k] : acc[k]) + b_dt * kk[0];  // the running sum starts from x
      if (stage < 4) {
        const double v = xs[k] + a_dt * kk[0];
        dst[q] = v;
        own[(stage + 1) & 1][k] = v;
      } else if (r0 + y < MM && c0 + tx < MM) {
        // new state of a real cell of the tile (cells of a ragged tile beyond the lattice are periodic images)
        a.x_out[(long long)(r0 + y) * MM + (c0 + tx)] = acc[k];
      }
    }
    if (stage < 4) {
#pragma unroll
      for (int e = 0; e < NF; ++e) {
        if (fd[e] == 0 || fd[e] > (4 - stage) * H) continue;
        const int w = fw[e];
        const int y = w / PW - mg, x = w - (y + mg) * PW - mg;
        at_cell(y, x);
        double kk[1] = {0.0};
        view.p = src + w;
        view.own = src[w];
        view.mine = nullptr;
        sys.cell(view, kk, idx, ts);
        dst[w] = fx[e] + a_dt * kk[0];
      }
      __syncthreads();
    }
  }
  if (view.bad) *a.flag = 1;
}

}  // namespace odeintr_b200

#if ODEINTR_LATTICE_STAGES == 4
extern "C" __global__ void __launch_bounds__(256, ODEINTR_TILE2_MIN_BLOCKS)
odeintr_lattice_rk4_tiled2d_h1(const odeintr_lattice_tiled_args a) { odeintr_b200::lattice_rk4_tiled2d_body<1>(a); }

extern "C" __global__ void __launch_bounds__(256, 2)
odeintr_lattice_rk4_tiled2d_h2(const odeintr_lattice_tiled_args a) { odeintr_b200::lattice_rk4_tiled2d_body<2>(a); }
#endif

#endif  // ODEINTR_LATTICE_2D
