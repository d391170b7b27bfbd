// odeintr_b200/csrc/kernels/lattice_dopri5_stream2d.cuh
// K4ds2: one ATTEMPTED Dormand-Prince 5(4) step of a periodic M x M lattice in one pass -- compile_sys_openmp's DEFAULT
// method (rk5_i; also rk5_a and plain rk5) on the reference's own documented large-system example, bistable_at() on an
// M x M lattice (R/odeintr.R:393-421) -- by the warp-autonomous STREAMING scheme of lattice_stream2d.cuh.
// A warp owns a strip of 64 columns (two per lane) and marches down its rows.  Row j of (x, k1) enters a pipeline of the
// six neighbour-dependent stages, all held in registers:
//     row j   : s2 = x + dt b21 k1                         (pointwise)
//     row j-1 : k2 = f(s2 rows j-2..j)      -> s3 of row j-1
//     row j-2 : k3 = f(s3 rows j-3..j-1)    -> s4 of row j-2
//     row j-3 : k4                          -> s5 of row j-3
//     row j-4 : k5                          -> s6 of row j-4
//     row j-5 : k6                          -> x_new of row j-5                  (stored)
//     row j-6 : k7 = f(x_new rows j-7..j-5) -> error estimate of row j-6, k7     (stored)
// The stage derivatives are never kept: the moment k_i of a row is known, its term is added to the PARTIAL SUMS of all
// later stage states of that row (s_{i+1} .. s6, x_new, the error estimate).  Boost's scale_sumN adds the terms left
// to right in the order k1, k2, ..., which is the order the k_i are born in, so every sum sees the same roundings:
// bit-identical to the stage-per-launch schedule.  A row carries x and k1 at age 0, then 5, 4, 3, 2, 1 values.
// Each stage keeps a window of three rows of the lane's two columns; the columns beside a lane's pair are fetched from
// the neighbouring lane by a shuffle when the right-hand side asks for them (lazy: the windows hold 2 doubles per row,
// not 4).  The row loop is unrolled by six (LCM of the window rotation 3 and the age rotation 6): every register slot
// is a compile-time index, nothing is shifted.  What the edge lanes lack spoils one more column per stage and side: a
// strip yields 52 of its 64 columns; 12 rows of pipeline fill per row segment.  HBM sees x and k1 once (x 64/52),
// x_new and k7 once; |x| and |k1| for the error norm are re-read when the row leaves the pipeline (an L1 / L2 hit).
// Dense output: the accepted attempt is repeated with the interpolation polynomial in place of the error estimate
// (lattice_dopri5_tiled.cuh), so k3 .. k6 never reach HBM.
// Radius 1, one field, even M >= 64, loops over all cells; everything else keeps the seven combo launches per attempt.
#pragma once
#include "odeintr_b200/lattice_lazy2_view.cuh"

#if defined(ODEINTR_LATTICE_2D) && ODEINTR_NFIELDS == 1 && !defined(ODEINTR_NO_SYMBOLIC) && ODEINTR_LATTICE_STAGES == 7

#ifndef ODEINTR_STREAM5_MIN_BLOCKS
#define ODEINTR_STREAM5_MIN_BLOCKS 1
#endif
#ifndef ODEINTR_STREAM5_THREADS  // threads per block (the host launches what odeintr_lattice_dopri5_stream2d_info reports)
#define ODEINTR_STREAM5_THREADS 256
#endif

namespace odeintr_b200 {

ODEINTR_DEV void lattice_dopri5_stream2d_body(const odeintr_lattice_dopri5_args& a) {
  typedef odeintr::system_type_t<lattice_lazy2_view> Sys;
  constexpr int MM = isqrt(ODEINTR_SYS_SIZE);
  constexpr int CW = 64, MARG = 6, USE = CW - 2 * MARG;  // columns of a strip, columns spoilt per side, columns it yields
  constexpr int PF = 2;                                  // rows prefetched ahead (an iteration of a warp outlasts the HBM latency)
  constexpr int n_strips = (MM + USE - 1) / USE;
  const int RSEG = a.seg_rows > 0 ? a.seg_rows : 256;
  const int n_segs = (MM + RSEG - 1) / RSEG;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const long long task0 = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_tasks = static_cast<long long>(n_strips) * n_segs;
  const long long task_stride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);

  Sys sys;
  sys.load_params(a.pars, 1, 0);
  lattice_lazy2_view view;
  view.side = MM; view.bad = 0;
  rel_index2 idx;
  idx.side = MM;
  const bool dense = a.dense_out != nullptr;
  // coefficient * dt products and stage times: formed once on the host (the same IEEE products), read here as
  // constant-bank operands -- two dozen doubles that would otherwise live in registers or be re-formed in the loop
  const double* const cf = a.cf;
  auto wrap = [](int v) {
    while (v < 0) v += MM;
    while (v >= MM) v -= MM;
    return v;
  };
  double emax = 0.0;
  enum { QX = 0, QK = 1, S4 = 2, S5 = 3, S6 = 4, SX = 5, SE = 6, NSUM = 7 };  // (x and k1 of a row travel until its sums have all started)

  for (long long task = task0; task < n_tasks; task += task_stride) {
    const int strip = static_cast<int>(task % n_strips), seg = static_cast<int>(task / n_strips);
    const int c_out = strip * USE;                 // first column the strip yields
    const int r_lo = seg * RSEG, r_hi = r_lo + RSEG < MM ? r_lo + RSEG : MM;
    const int col0 = wrap(c_out - MARG + 2 * lane);  // the lane's first column (even: M is even), its pair does not straddle the seam
    const bool keeps = lane >= MARG / 2 && lane < 32 - MARG / 2 && c_out + 2 * (lane - MARG / 2) < MM;

    // windows of the six stage inputs (s2 .. s6, x_new): three rows of the lane's pair each, rotating through three
    // register rows; partial sums of the rows in flight, rotating through six
    double W[6][3][2];
    double Q[6][NSUM][2];
#pragma unroll
    for (int s = 0; s < 6; ++s)
#pragma unroll
      for (int r = 0; r < 3; ++r) W[s][r][0] = W[s][r][1] = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
      for (int s = 0; s < NSUM; ++s) Q[q][s][0] = Q[q][s][1] = 0.0;

    double2 pfx[PF], pfk[PF];
    const int j0 = r_lo - MARG;
    const int j1 = j0 + (r_hi + MARG - j0 + 5) / 6 * 6;   // rows [j0, j1) enter the pipeline: a multiple of six
#pragma unroll
    for (int p = 0; p < PF; ++p) {
      const long long o = static_cast<long long>(wrap(j0 + p)) * MM + col0;
      pfx[p] = __ldg(reinterpret_cast<const double2*>(a.x_in + o));
      pfk[p] = __ldg(reinterpret_cast<const double2*>(a.d_in + o));
    }

    for (int jb = j0; jb < j1; jb += 6) {
#pragma unroll
      for (int ph = 0; ph < 6; ++ph) {
        const int j = jb + ph;
        const int nw = ph % 3, md = (ph + 2) % 3, ol = (ph + 1) % 3;  // window slots of the newest / middle / oldest row
        // k of the lane's two cells in the middle row of window s (global row `row`)
        auto rhs = [&](const int s, const int row, const double ts, double& k0, double& k1v) {
          const int rw = wrap(row);
          double kk[1];
          view.up = W[s][ol]; view.mid = W[s][md]; view.dn = W[s][nw];
          view.left = __shfl_up_sync(full, W[s][md][1], 1);     // the left neighbour's second column
          view.right = __shfl_down_sync(full, W[s][md][0], 1);  // the right neighbour's first column
          view.row = idx.row = rw;
          kk[0] = 0.0;
          view.ci = 0; view.col = idx.col = col0; idx.centre = rw * MM + col0;
          sys.cell(view, kk, idx, ts);
          k0 = kk[0];
          kk[0] = 0.0;
          view.ci = 1; view.col = idx.col = col0 + 1; idx.centre = rw * MM + col0 + 1;
          sys.cell(view, kk, idx, ts);
          k1v = kk[0];
        };
        const double2 xin = pfx[ph % PF], kin = pfk[ph % PF];
        if (j + PF < j1) {
          const long long o = static_cast<long long>(wrap(j + PF)) * MM + col0;
          pfx[ph % PF] = __ldg(reinterpret_cast<const double2*>(a.x_in + o));
          pfk[ph % PF] = __ldg(reinterpret_cast<const double2*>(a.d_in + o));
        }
        // |x| and |k1| of the row that leaves the pipeline in this iteration (error norm): asked for now, used at the end
        double2 xo, ko;
        {
          const int ro = j - 6;
          const bool want = !dense && keeps && ro >= r_lo && ro < r_hi;
          const long long o = static_cast<long long>(want ? ro : r_lo) * MM + col0;
#ifndef ODEINTR_STREAM5_LATE_NORM
          xo = __ldg(reinterpret_cast<const double2*>(a.x_in + o));
          ko = __ldg(reinterpret_cast<const double2*>(a.d_in + o));
#endif
        }
        double k0, k1v;
        // s2 of row j (pointwise)
        W[0][nw][0] = xin.x + cf[ODEINTR_CF_B21] * kin.x; W[0][nw][1] = xin.y + cf[ODEINTR_CF_B21] * kin.y;
        {  // k2 at row j-1 -> s3; the sums k2 contributes to start here, from x and k1 of the row
          double (&q)[NSUM][2] = Q[(ph + 5) % 6];
          rhs(0, j - 1, cf[ODEINTR_CF_T2], k0, k1v);
          const double x0 = q[QX][0], x1 = q[QX][1], d0 = q[QK][0], d1 = q[QK][1];
          double v0, v1;
          v0 = x0 + cf[ODEINTR_CF_B31] * d0; v1 = x1 + cf[ODEINTR_CF_B31] * d1;
          W[1][nw][0] = v0 + cf[ODEINTR_CF_B32] * k0; W[1][nw][1] = v1 + cf[ODEINTR_CF_B32] * k1v;
          v0 = x0 + cf[ODEINTR_CF_B41] * d0; v1 = x1 + cf[ODEINTR_CF_B41] * d1;
          q[S4][0] = v0 + cf[ODEINTR_CF_B42] * k0; q[S4][1] = v1 + cf[ODEINTR_CF_B42] * k1v;
          v0 = x0 + cf[ODEINTR_CF_B51] * d0; v1 = x1 + cf[ODEINTR_CF_B51] * d1;
          q[S5][0] = v0 + cf[ODEINTR_CF_B52] * k0; q[S5][1] = v1 + cf[ODEINTR_CF_B52] * k1v;
          v0 = x0 + cf[ODEINTR_CF_B61] * d0; v1 = x1 + cf[ODEINTR_CF_B61] * d1;
          q[S6][0] = v0 + cf[ODEINTR_CF_B62] * k0; q[S6][1] = v1 + cf[ODEINTR_CF_B62] * k1v;
        }
        {  // k3 at row j-2 -> s4; x_new and the error estimate (dense-output polynomial) start here
          double (&q)[NSUM][2] = Q[(ph + 4) % 6];
          rhs(1, j - 2, cf[ODEINTR_CF_T3], k0, k1v);
          const double x0 = q[QX][0], x1 = q[QX][1], d0 = q[QK][0], d1 = q[QK][1];
          W[2][nw][0] = q[S4][0] + cf[ODEINTR_CF_B43] * k0; W[2][nw][1] = q[S4][1] + cf[ODEINTR_CF_B43] * k1v;
          q[S5][0] = q[S5][0] + cf[ODEINTR_CF_B53] * k0; q[S5][1] = q[S5][1] + cf[ODEINTR_CF_B53] * k1v;
          q[S6][0] = q[S6][0] + cf[ODEINTR_CF_B63] * k0; q[S6][1] = q[S6][1] + cf[ODEINTR_CF_B63] * k1v;
          double v0 = x0 + cf[ODEINTR_CF_C1] * d0, v1 = x1 + cf[ODEINTR_CF_C1] * d1;
          q[SX][0] = v0 + cf[ODEINTR_CF_C3] * k0; q[SX][1] = v1 + cf[ODEINTR_CF_C3] * k1v;
          if (dense) { v0 = x0 + cf[ODEINTR_CF_E1] * d0; v1 = x1 + cf[ODEINTR_CF_E1] * d1; }
          else { v0 = cf[ODEINTR_CF_E1] * d0; v1 = cf[ODEINTR_CF_E1] * d1; }
          q[SE][0] = v0 + cf[ODEINTR_CF_E3] * k0; q[SE][1] = v1 + cf[ODEINTR_CF_E3] * k1v;
        }
        {  // k4 at row j-3 -> s5
          double (&q)[NSUM][2] = Q[(ph + 3) % 6];
          rhs(2, j - 3, cf[ODEINTR_CF_T4], k0, k1v);
          W[3][nw][0] = q[S5][0] + cf[ODEINTR_CF_B54] * k0; W[3][nw][1] = q[S5][1] + cf[ODEINTR_CF_B54] * k1v;
          q[S6][0] = q[S6][0] + cf[ODEINTR_CF_B64] * k0; q[S6][1] = q[S6][1] + cf[ODEINTR_CF_B64] * k1v;
          q[SX][0] = q[SX][0] + cf[ODEINTR_CF_C4] * k0; q[SX][1] = q[SX][1] + cf[ODEINTR_CF_C4] * k1v;
          q[SE][0] = q[SE][0] + cf[ODEINTR_CF_E4] * k0; q[SE][1] = q[SE][1] + cf[ODEINTR_CF_E4] * k1v;
        }
        {  // k5 at row j-4 -> s6
          double (&q)[NSUM][2] = Q[(ph + 2) % 6];
          rhs(3, j - 4, cf[ODEINTR_CF_T5], k0, k1v);
          W[4][nw][0] = q[S6][0] + cf[ODEINTR_CF_B65] * k0; W[4][nw][1] = q[S6][1] + cf[ODEINTR_CF_B65] * k1v;
          q[SX][0] = q[SX][0] + cf[ODEINTR_CF_C5] * k0; q[SX][1] = q[SX][1] + cf[ODEINTR_CF_C5] * k1v;
          q[SE][0] = q[SE][0] + cf[ODEINTR_CF_E5] * k0; q[SE][1] = q[SE][1] + cf[ODEINTR_CF_E5] * k1v;
        }
        {  // k6 at row j-5 -> x_new
          double (&q)[NSUM][2] = Q[(ph + 1) % 6];
          rhs(4, j - 5, cf[ODEINTR_CF_T6], k0, k1v);
          const double n0 = q[SX][0] + cf[ODEINTR_CF_C6] * k0, n1 = q[SX][1] + cf[ODEINTR_CF_C6] * k1v;
          W[5][nw][0] = n0; W[5][nw][1] = n1;
          q[SE][0] = q[SE][0] + cf[ODEINTR_CF_E6] * k0; q[SE][1] = q[SE][1] + cf[ODEINTR_CF_E6] * k1v;
          const int ro = j - 5;
          if (!dense && keeps && ro >= r_lo && ro < r_hi) {
            double2 o;
            o.x = n0; o.y = n1;
            *reinterpret_cast<double2*>(a.x_out + static_cast<long long>(ro) * MM + col0) = o;
          }
        }
        {  // k7 = f(x_new) at row j-6: the derivative handed to the next step, the last term of the error estimate
          double (&q)[NSUM][2] = Q[ph];
          rhs(5, j - 6, cf[ODEINTR_CF_T6], k0, k1v);
          const double v0 = q[SE][0] + cf[ODEINTR_CF_E7] * k0, v1 = q[SE][1] + cf[ODEINTR_CF_E7] * k1v;
          const int ro = j - 6;
          if (keeps && ro >= r_lo && ro < r_hi) {
            const long long o = static_cast<long long>(ro) * MM + col0;
            double2 w;
            if (dense) {
              w.x = v0; w.y = v1;
              *reinterpret_cast<double2*>(a.dense_out + o) = w;
            } else {
              w.x = k0; w.y = k1v;
              *reinterpret_cast<double2*>(a.d_out + o) = w;
#ifdef ODEINTR_STREAM5_LATE_NORM
              xo = __ldg(reinterpret_cast<const double2*>(a.x_in + o));
              ko = __ldg(reinterpret_cast<const double2*>(a.d_in + o));
#endif
              const double ea = fabs(v0) / (a.eps_abs + a.eps_rel * (1.0 * fabs(xo.x) + a.adt * fabs(ko.x)));
              const double eb = fabs(v1) / (a.eps_abs + a.eps_rel * (1.0 * fabs(xo.y) + a.adt * fabs(ko.y)));
              emax = fmax(emax, fmax(ea, eb));
            }
          }
        }
        {  // x and k1 of row j travel with it (its slot was row j-6's until a moment ago)
          double (&q)[NSUM][2] = Q[ph];
          q[QX][0] = xin.x; q[QX][1] = xin.y;
          q[QK][0] = kin.x; q[QK][1] = kin.y;
        }
      }
    }
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) emax = fmax(emax, __shfl_xor_sync(full, emax, sft));
  if (lane == 0 && emax > 0.0) atomicMax(a.err_max, (unsigned long long)__double_as_longlong(emax));
  if (view.bad) *a.flag = 1;
}

}  // namespace odeintr_b200

extern "C" __global__ void __launch_bounds__(ODEINTR_STREAM5_THREADS, ODEINTR_STREAM5_MIN_BLOCKS)
odeintr_lattice_dopri5_stream2d(const odeintr_lattice_dopri5_args a) { odeintr_b200::lattice_dopri5_stream2d_body(a); }

extern "C" __global__ void odeintr_lattice_dopri5_stream2d_info(int* out) { out[0] = ODEINTR_STREAM5_THREADS; }

#endif  // dopri5 on an M x M lattice
