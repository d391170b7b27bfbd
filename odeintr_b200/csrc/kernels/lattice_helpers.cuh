// odeintr_b200/csrc/kernels/lattice_helpers.cuh
// Names a large-system snippet may call (the reference offers them through inst/include/utils.h:6-31):
// M, bound(), d_left/d_up/d_right/d_down and laplace2D on a periodic M x M row-major lattice.
// This fragment is included INSIDE odeintr::system_type, where N and the state view are in scope.
// S is any indexable state view (register vector, global-memory view or shared-memory tile view).
static constexpr int M = odeintr_b200::isqrt(N);

// periodic wrap of an arbitrary (possibly negative) coordinate into [0, M)
__device__ __forceinline__ static int bound(const int c) {
  const int r = c % M;
  return r + (r < 0 ? M : 0);
}

#define ODEINTR_NEIGHBOUR_DIFF(NAME, ROW, COL)                                        \
  template <class S>                                                                  \
  __device__ __forceinline__ static double NAME(const S& x, const int i, const int j) { \
    return x[(ROW) * M + (COL)] - x[i * M + j];                                       \
  }
ODEINTR_NEIGHBOUR_DIFF(d_left, i, bound(j - 1))
ODEINTR_NEIGHBOUR_DIFF(d_up, bound(i - 1), j)
ODEINTR_NEIGHBOUR_DIFF(d_right, i, bound(j + 1))
ODEINTR_NEIGHBOUR_DIFF(d_down, bound(i + 1), j)
#undef ODEINTR_NEIGHBOUR_DIFF

// 5-point periodic Laplacian, summed in the reference's order: left + up + right + down
template <class S>
__device__ __forceinline__ static double laplace2D(const S& x, const int i, const int j) {
  return d_left(x, i, j) + d_up(x, i, j) + d_right(x, i, j) + d_down(x, i, j);
}
