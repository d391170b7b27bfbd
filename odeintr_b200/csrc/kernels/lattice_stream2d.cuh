// odeintr_b200/csrc/kernels/lattice_stream2d.cuh
// K4s2: a whole rk4 step of a periodic M x M lattice (compile_sys_openmp's documented example, laplace2D / laplace4,
// /root/reference R/odeintr.R:393-413) by warp-autonomous STREAMING: the temporally blocked step of
// lattice_tiled2d.cuh without shared memory, without barriers and without redundant rows.
//
// A warp owns a strip of 64 columns (two per lane) and marches down its rows.  Row j of x enters at the top of a
// four-stage pipeline held in registers:
//     row j   : x                       (loaded)
//     row j-1 : k1 = f(x rows j-2..j)            -> stage-1 state, running sum
//     row j-2 : k2 = f(stage-1 rows j-3..j-1)    -> stage-2 state, running sum
//     row j-3 : k3 = f(stage-2 rows j-4..j-2)    -> stage-3 state, running sum
//     row j-4 : k4 = f(stage-3 rows j-5..j-3)    -> new state of row j-4 (stored)
// so every loaded row yields one finished row: the rows above / below a cell are the lane's own registers (each stage
// keeps a window of three rows), the columns left / right of a lane's pair come from the neighbouring lanes by
// shuffles -- once per row and stage, kept with the row.  Only the strip's outer 4 columns on either side are
// redundant (a strip yields 56 of its 64 columns: what the edge lanes lack spoils one more column per stage), plus 8
// rows of pipeline fill per row segment.  HBM sees every cell once for reading (x 64/56) and once for writing, as
// coalesced 16-byte accesses per lane; the next rows are prefetched into registers.
// Rows and columns wrap periodically in the address arithmetic, so every strip is treated alike.
// Radius 1 (the 5- and 9-point stencils), loops over all cells, even M; everything else stays with K4t2.
// Every cell goes through exactly the operations of the staged schedule (Boost's order, no FMA): bit-identical.
// Measured variants (profiles/r02_stream2d.md): 2 blocks of 128 registers per SM 1.50-1.53e11 cell-steps/s at 8192^2
// (K4t2: 1.30e11), 1 block of 138 registers 1.07-1.09e11; stages two rows apart (four independent right-hand sides per
// iteration, 196 registers, one block per SM) 1.06e11 with long_scoreboard the top stall -- the wider live set costs
// more in occupancy than the independence gains; with a 128-register cap it spills (5.7e10).  Kept: this form.
#pragma once
#include "odeintr_b200/lattice_tiled2d.cuh"

#if defined(ODEINTR_LATTICE_2D) && ODEINTR_NFIELDS == 1 && !defined(ODEINTR_NO_SYMBOLIC) && ODEINTR_LATTICE_STAGES == 4

#ifndef ODEINTR_STREAM2_ROWS  // rows of a segment (one warp task): 8 more are computed to fill the pipeline
#define ODEINTR_STREAM2_ROWS 512
#endif
#ifndef ODEINTR_STREAM2_MIN_BLOCKS  // measured (profiles/r02_stream2d.md): 2 blocks of 128 registers beat 1 block of 138
#define ODEINTR_STREAM2_MIN_BLOCKS 2
#endif
#ifndef ODEINTR_STREAM2_L2_AHEAD  // rows ahead of the register prefetch queue that are pulled into L2 (measured: -3 %, off)
#define ODEINTR_STREAM2_L2_AHEAD 0
#endif

namespace odeintr_b200 {

// One stage's input around the cell being computed: three rows (above, own, below) of the lane's two columns plus the
// column on either side, in registers.  `up` / `mid` / `dn` point at the cell's entry of the row above / its own row /
// the row below (the three rows of a window rotate through three register rows: no row is ever moved).
struct lattice_reg2_view {
  static constexpr int RS = 4;  // left neighbour, the lane's two columns, right neighbour
  const double *up, *mid, *dn;
  int row, col, side;
  mutable int bad;

  ODEINTR_DEV double at2(long long dr, long long dc) const {
    if (dr > 1) { bad = 1; dr = 1; }
    if (dr < -1) { bad = 1; dr = -1; }
    if (dc > 1) { bad = 1; dc = 1; }
    if (dc < -1) { bad = 1; dc = -1; }
    double v = 0.0;
#pragma unroll
    for (int r = -1; r <= 1; ++r)
#pragma unroll
      for (int c = -1; c <= 1; ++c)
        if (dr == r && dc == c) v = (r < 0 ? up : r > 0 ? dn : mid)[c];  // never a run-time register index
    return v;
  }
  ODEINTR_DEV double operator[](const rel_cell2& c) const {
    if (c.rel && c.has_row && c.has_col) return at2(c.dr, c.dc);
    return (*this)[c.abs];
  }
  ODEINTR_DEV double operator[](const rel_index2&) const { return mid[0]; }
  ODEINTR_DEV double operator[](const long long g) const {
    if (g < 0 || g >= static_cast<long long>(side) * side) { bad = 1; return mid[0]; }
    const long long r = g / side;
    return at2(min_image(r - row, side), min_image(g - r * side - col, side));
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(side) * side; }
};

typedef lattice_reg2_view stream2_view;  // (measured: the lazy two-doubles-per-row view of K4ds2 loses here, 1.66e11 vs 1.75e11)

ODEINTR_DEV void lattice_rk4_stream2d_body(const odeintr_lattice_tiled_args& a) {
  typedef odeintr::system_type_t<stream2_view> Sys;
  constexpr int MM = isqrt(ODEINTR_SYS_SIZE);
  constexpr int CW = 64, USE = CW - 8;      // columns of a strip, columns it yields
  constexpr int RSEG = ODEINTR_STREAM2_ROWS;
  constexpr int RS = lattice_reg2_view::RS, C0 = 1;  // a window row: left neighbour, the two columns, right neighbour
  constexpr int PF = 3;                     // rows prefetched ahead
  constexpr int n_strips = (MM + USE - 1) / USE, n_segs = (MM + RSEG - 1) / RSEG;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const long long task0 = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_tasks = static_cast<long long>(n_strips) * n_segs;
  const long long task_stride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);

  Sys sys;
  sys.load_params(a.pars, 1, 0);
  stream2_view view;
  view.side = MM; view.bad = 0;
  rel_index2 idx;
  idx.side = MM;
  const double half = 1.0 / 2.0, b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
  const double dt = a.dt;
  const double t1 = a.t, t2 = a.t + half * dt, t4 = a.t + 1.0 * dt;
  auto wrap = [](int v) {
    while (v < 0) v += MM;
    while (v >= MM) v -= MM;
    return v;
  };

  for (long long task = task0; task < n_tasks; task += task_stride) {
    const int strip = static_cast<int>(task % n_strips), seg = static_cast<int>(task / n_strips);
    const int c_out = strip * USE;                 // first column the strip yields
    const int r_lo = seg * RSEG, r_hi = r_lo + RSEG < MM ? r_lo + RSEG : MM;
    const int col0 = wrap(c_out - 4 + 2 * lane);   // the lane's first column (even: M is even), its pair does not straddle the seam
    const bool keeps = lane >= 2 && lane < 30 && c_out + 2 * (lane - 2) < MM;  // the lane's pair belongs to the output
    const int col_left = wrap(col0 - 1), col_right = wrap(col0 + 2);

    // windows of the four stage inputs: three rows each, [left, c0, c1, right].  The rows ROTATE through the three
    // register rows (the row loop is unrolled by three with compile-time slots), so a new row overwrites the oldest
    // one and nothing is shifted: the row entering in phase ph lives in slot ph, the one above it in slot (ph + 2) % 3
    double W[4][3][RS];
    double x3[2], acc1[2], acc2[2], acc3[2];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < RS; ++q) W[s][r][q] = 0.0;
    x3[0] = x3[1] = acc1[0] = acc1[1] = acc2[0] = acc2[1] = acc3[0] = acc3[1] = 0.0;

    // prefetch queue of input rows (slot = phase)
    double2 pf[PF];
    const int j0 = r_lo - 4;
    const int j1 = j0 + (r_hi + 4 - j0 + 2) / 3 * 3;   // rows [j0, j1) enter the pipeline: a multiple of three
#pragma unroll
    for (int p = 0; p < PF; ++p)
      pf[p] = __ldg(reinterpret_cast<const double2*>(a.x_in + static_cast<long long>(wrap(j0 + p)) * MM + col0));

    for (int jb = j0; jb < j1; jb += 3) {
#pragma unroll
      for (int ph = 0; ph < 3; ++ph) {
        const int j = jb + ph;
        const int nw = ph, md = (ph + 2) % 3, ol = (ph + 1) % 3;  // slots of the newest / middle / oldest row after a push
        // a new row enters stage window s (over the row that leaves it); the neighbouring lanes' edge columns with it
        auto push = [&](const int s, const double v0, const double v1) {
          W[s][nw][C0] = v0; W[s][nw][C0 + 1] = v1;
          W[s][nw][0] = __shfl_up_sync(full, v1, 1);    // the left neighbour's second column
          W[s][nw][3] = __shfl_down_sync(full, v0, 1);  // the right neighbour's first column
        };
        // k of the lane's two cells in the middle row of window s (global row `row`)
        auto rhs = [&](const int s, const int row, const double ts, double& k0, double& k1v) {
          const int rw = wrap(row);
          double kk[1];
          kk[0] = 0.0;
          view.up = &W[s][ol][1]; view.mid = &W[s][md][1]; view.dn = &W[s][nw][1];
          view.row = idx.row = rw; view.col = idx.col = col0; idx.centre = rw * MM + col0;
          sys.cell(view, kk, idx, ts);
          k0 = kk[0];
          kk[0] = 0.0;
          view.up = &W[s][ol][2]; view.mid = &W[s][md][2]; view.dn = &W[s][nw][2];
          view.col = idx.col = col0 + 1; idx.centre = rw * MM + col0 + 1;
          sys.cell(view, kk, idx, ts);
          k1v = kk[0];
        };

        const double2 xin = pf[ph];
        // (measured: loading unconditionally -- rows wrap, the rows past the segment's end exist -- removes the
        // predicated copy that ncu shows waiting for the load, but costs a spill at 128 registers: 1.72e11 vs 1.75e11)
        if (j + PF < j1)
          pf[ph] = __ldg(reinterpret_cast<const double2*>(a.x_in + static_cast<long long>(wrap(j + PF)) * MM + col0));
        if (ODEINTR_STREAM2_L2_AHEAD > 0 && (lane & 7) == 0 && j + PF + ODEINTR_STREAM2_L2_AHEAD < j1)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x_in + static_cast<long long>(wrap(j + PF + ODEINTR_STREAM2_L2_AHEAD)) * MM + col0));

        // x of row j - 3 leaves the x window with this push: stage 3 needs it
        x3[0] = W[0][nw][C0]; x3[1] = W[0][nw][C0 + 1];
        push(0, xin.x, xin.y);                       // x rows j-2, j-1, j
        double k0, k1v;
        // stage 1 at row j-1
        rhs(0, j - 1, t1, k0, k1v);
        const double s1a = W[0][md][C0] + (half * dt) * k0, s1b = W[0][md][C0 + 1] + (half * dt) * k1v;
        const double n1a = W[0][md][C0] + (b1 * dt) * k0, n1b = W[0][md][C0 + 1] + (b1 * dt) * k1v;   // the running sum starts from x
        push(1, s1a, s1b);                           // stage-1 states of rows j-3, j-2, j-1
        // stage 2 at row j-2
        rhs(1, j - 2, t2, k0, k1v);
        const double s2a = W[0][ol][C0] + (half * dt) * k0, s2b = W[0][ol][C0 + 1] + (half * dt) * k1v;
        const double n2a = acc1[0] + (b2 * dt) * k0, n2b = acc1[1] + (b2 * dt) * k1v;
        push(2, s2a, s2b);                           // stage-2 states of rows j-4, j-3, j-2
        // stage 3 at row j-3
        rhs(2, j - 3, t2, k0, k1v);
        const double s3a = x3[0] + (1.0 * dt) * k0, s3b = x3[1] + (1.0 * dt) * k1v;
        const double n3a = acc2[0] + (b2 * dt) * k0, n3b = acc2[1] + (b2 * dt) * k1v;
        push(3, s3a, s3b);                           // stage-3 states of rows j-5, j-4, j-3
        // stage 4 at row j-4: the new state
        rhs(3, j - 4, t4, k0, k1v);
        const double o0 = acc3[0] + (b1 * dt) * k0, o1 = acc3[1] + (b1 * dt) * k1v;
        const int ro = j - 4;
        if (keeps && ro >= r_lo && ro < r_hi) {
          double2 o;
          o.x = o0; o.y = o1;
          *reinterpret_cast<double2*>(a.x_out + static_cast<long long>(ro) * MM + col0) = o;
        }
        acc3[0] = n3a; acc3[1] = n3b;
        acc2[0] = n2a; acc2[1] = n2b;
        acc1[0] = n1a; acc1[1] = n1b;
      }
    }
    (void)col_left; (void)col_right;
  }
  if (view.bad) *a.flag = 1;
}

}  // namespace odeintr_b200

extern "C" __global__ void __launch_bounds__(256, ODEINTR_STREAM2_MIN_BLOCKS)
odeintr_lattice_rk4_stream2d(const odeintr_lattice_tiled_args a) { odeintr_b200::lattice_rk4_stream2d_body(a); }

// rows per segment the kernel was compiled with (the host sizes the grid from it)
extern "C" __global__ void odeintr_lattice_stream2d_info(int* out) { out[0] = ODEINTR_STREAM2_ROWS; }

#endif  // rk4 on an M x M lattice
