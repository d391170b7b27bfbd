// odeintr_b200/csrc/kernels/lattice_lazy2_view.cuh
// The state view of the streaming M x M kernels (lattice_stream2d.cuh, lattice_dopri5_stream2d.cuh).
#pragma once
#include "odeintr_b200/lattice_tiled2d.cuh"

#if defined(ODEINTR_LATTICE_2D) && ODEINTR_NFIELDS == 1 && !defined(ODEINTR_NO_SYMBOLIC)
namespace odeintr_b200 {

// One stage's input around the cell being computed: the lane's two columns of three rows, in registers.  The columns
// beside the pair live in the neighbouring lanes: those of the cell's own row are fetched by two shuffles before the
// right-hand side is evaluated (every lane, straight-line code); those of the rows above and below -- the diagonal
// neighbours of a 9-point stencil -- by a shuffle on demand.  All lanes evaluate the same expression at the same time
// (the displacements are the snippet's constants), so those shuffles are convergent too; __activemask() keeps a snippet
// with lane-dependent displacements from hanging -- it is flagged and the call falls back to the staged kernels.
struct lattice_lazy2_view {
  const double *up, *mid, *dn;  // the pair (c0, c1) of the row above / the cell's own row / the row below
  double left, right;            // the columns beside the pair in the cell's own row
  int ci;                        // which column of the pair is being computed
  int row, col, side;
  mutable int bad;

  ODEINTR_DEV double fetch(const int r, const int p) const {
    const double* q = r < 0 ? up : r > 0 ? dn : mid;
    if (p == 0) return q[0];
    if (p == 1) return q[1];
    if (r == 0) return p < 0 ? left : right;
    const unsigned m = __activemask();
    bad |= (m != 0xffffffffu);
    if (p < 0) return __shfl_up_sync(m, q[1], 1);  // the left neighbour's second column
    return __shfl_down_sync(m, q[0], 1);           // the right neighbour's first column
  }
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
        if (dr == r && dc == c) v = fetch(r, ci + c);  // never a run-time register index
    return v;
  }
  ODEINTR_DEV double operator[](const rel_cell2& c) const {
    if (c.rel && c.has_row && c.has_col) return at2(c.dr, c.dc);
    return (*this)[c.abs];
  }
  ODEINTR_DEV double operator[](const rel_index2&) const { return mid[ci]; }
  // an absolute flat index (hand-written index arithmetic): nearest periodic image of the cell it points at
  ODEINTR_DEV double operator[](const long long g) const {
    if (g < 0 || g >= static_cast<long long>(side) * side) { bad = 1; return mid[ci]; }
    const long long r = g / side;
    return at2(min_image(r - row, side), min_image(g - r * side - col, side));
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(side) * side; }
};

}  // namespace odeintr_b200
#endif
