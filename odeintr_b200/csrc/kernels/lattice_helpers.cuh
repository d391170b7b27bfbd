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

#ifdef ODEINTR_HAVE_REL_INDEX
// symbolic coordinates (lattice_index.cuh): the wrapped neighbour is the nearest periodic image, its displacement stays
__device__ __forceinline__ static odeintr_b200::rel_coord bound(odeintr_b200::rel_coord c) {
  if (c.rel) c.wrapped = true;
  else c.off = bound(static_cast<int>(c.off));
  return c;
}
#define ODEINTR_REL_COORD_OVERLOAD(NAME, ROW, COL)                                    \
  template <class S>                                                                  \
  __device__ __forceinline__ static double NAME(const S& x, const odeintr_b200::rel_coord i,  \
                                                const odeintr_b200::rel_coord j) {    \
    return x[(ROW) * M + (COL)] - x[i * M + j];                                       \
  }
#else
#define ODEINTR_REL_COORD_OVERLOAD(NAME, ROW, COL)
#endif

// the same expressions for plain coordinates (staged kernels) and symbolic ones (2-D tiled kernel)
#define ODEINTR_NEIGHBOUR_DIFF(NAME, ROW, COL)                                        \
  template <class S>                                                                  \
  __device__ __forceinline__ static double NAME(const S& x, const int i, const int j) { \
    return x[(ROW) * M + (COL)] - x[i * M + j];                                       \
  }                                                                                   \
  ODEINTR_REL_COORD_OVERLOAD(NAME, ROW, COL)
ODEINTR_NEIGHBOUR_DIFF(d_left, i, bound(j - 1))
ODEINTR_NEIGHBOUR_DIFF(d_up, bound(i - 1), j)
ODEINTR_NEIGHBOUR_DIFF(d_right, i, bound(j + 1))
ODEINTR_NEIGHBOUR_DIFF(d_down, bound(i + 1), j)
#undef ODEINTR_NEIGHBOUR_DIFF
#undef ODEINTR_REL_COORD_OVERLOAD

// 5-point periodic Laplacian, summed in the reference's order: left + up + right + down
template <class S>
__device__ __forceinline__ static double laplace2D(const S& x, const int i, const int j) {
  return d_left(x, i, j) + d_up(x, i, j) + d_right(x, i, j) + d_down(x, i, j);
}
#ifdef ODEINTR_HAVE_REL_INDEX
template <class S>
__device__ __forceinline__ static double laplace2D(const S& x, const odeintr_b200::rel_coord i,
                                                   const odeintr_b200::rel_coord j) {
  return d_left(x, i, j) + d_up(x, i, j) + d_right(x, i, j) + d_down(x, i, j);
}
#endif
