// odeintr_b200/csrc/kernels/lattice_view.cuh
// Read-only view of a large state vector in HBM, indexed by GLOBAL state index exactly as the user's
// snippet indexes the reference's std::vector (x[i - 1], x[(i + 1) % N], laplace2D(x, i, j), ...).
// On one GPU it is a plain pointer (SHARDED = false: no index arithmetic at all).  On a slab-sharded run
// (one process per GPU) the local buffer holds, per field, the rank's owned cells plus
// `ghost` cells on each side; a global index is wrapped periodically into that window.
#pragma once
#include "odeintr_b200/prelude.cuh"

namespace odeintr_b200 {

// MODE 0: plain pointer (one GPU, or an interior chunk of a one-field slab on a shifted base)
// MODE 1: slab window with periodic wrap (the chunks at the ring's seam)
// MODE 2: interior chunk of a slab with several fields: no wrap, but every field has its own constant shift
enum { LATTICE_PLAIN = 0, LATTICE_WRAP = 1, LATTICE_FIELDS = 2 };

template <int MODE>
struct lattice_view_t {
  const double* p;       // local cell 0 of field 0 (ghost cells sit at negative offsets)
  long long n_total;     // global cells per field (L)
  long long n_local;     // cells per field owned by this rank
  long long c0;          // global index of local cell 0
  long long ghost;       // ghost cells on each side
  long long pitch;       // distance between the local slabs of consecutive fields

  ODEINTR_DEV double operator[](const long long g) const {
    if (MODE == LATTICE_PLAIN) return __ldg(p + g);
#if defined(ODEINTR_NFIELDS) && ODEINTR_NFIELDS > 1
    constexpr int NF = ODEINTR_NFIELDS;
#else
    constexpr int NF = 1;
#endif
    if (MODE == LATTICE_FIELDS) {
      // entry (f, c) lives at f * pitch + (c - c0): a shift of f * (pitch - L) - c0 from its global index
      long long shift = -c0;
#pragma unroll
      for (int f = 1; f < NF; ++f)
        if (g >= f * n_total) shift = f * (pitch - n_total) - c0;
      return __ldg(p + g + shift);
    }
    // field and cell of the global index, then the periodic distance from the slab origin
    long long f = 0;
#pragma unroll
    for (int j = 1; j < NF; ++j)
      if (g >= j * n_total) f = j;
    long long d = (g - f * n_total) - c0;
    if (d < -ghost) d += n_total;
    else if (d >= n_local + ghost) d -= n_total;
    return __ldg(p + f * pitch + d);
  }
  ODEINTR_DEV std::size_t size() const { return static_cast<std::size_t>(n_total); }
};

}  // namespace odeintr_b200
