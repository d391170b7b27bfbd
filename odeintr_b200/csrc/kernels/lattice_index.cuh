// odeintr_b200/csrc/kernels/lattice_index.cuh
// The loop variable of a large-system snippet, as the kernels hand it to the user's loop body.
//
// In the reference the body of `for (int i = 0; i < N; ++i) dxdt[i] = ... x[(i + N - 1) % N] ...`
// (/root/reference R/odeintr.R:396-401) sees a plain int.  The staged kernels (lattice_rk.cuh, lattice_combo.cuh)
// keep that: loop_index<int>(c) is static_cast<int>(c).  The temporally blocked kernel (lattice_tiled.cuh) instead
// passes a rel_index: "the cell being computed plus a displacement", which follows +, - and % symbolically.  After
// inlining, the displacement of every x[...] of an ordinary stencil is a compile-time constant, so
// `x[(i + N - 1) % N]` becomes one shared-memory load at a fixed offset from the thread's own cell -- no modulo, no
// periodic-wrap selects, no 64-bit index arithmetic in the inner loop.  Anything the symbolic form cannot follow
// (i * 2, i / M, a modulo by something other than the lattice length, ...) degrades to the absolute index through
// the implicit conversion and takes the view's generic path: slower, same result.
#pragma once
#include "odeintr_b200/prelude.cuh"

namespace odeintr_b200 {

template <class T> struct is_integer { static const bool value = false; };
#define ODEINTR_INTEGER_(T) template <> struct is_integer<T> { static const bool value = true; };
ODEINTR_INTEGER_(char) ODEINTR_INTEGER_(signed char) ODEINTR_INTEGER_(unsigned char)
ODEINTR_INTEGER_(short) ODEINTR_INTEGER_(unsigned short) ODEINTR_INTEGER_(int) ODEINTR_INTEGER_(unsigned int)
ODEINTR_INTEGER_(long) ODEINTR_INTEGER_(unsigned long) ODEINTR_INTEGER_(long long) ODEINTR_INTEGER_(unsigned long long)
#undef ODEINTR_INTEGER_
template <bool B, class T> struct enable_if_ {};
template <class T> struct enable_if_<true, T> { typedef T type; };

struct rel_index {
  long long off;  // rel: displacement in index space from entry (field 0, centre); otherwise the absolute index
  int centre;     // global number of the cell being computed
  int cells;      // cells per field: the modulus of the periodic wrap
  bool rel;

  ODEINTR_DEV long long value() const { return rel ? static_cast<long long>(centre) + off : off; }
  ODEINTR_DEV operator long long() const { return value(); }
  ODEINTR_DEV rel_index absolute(const long long v) const {
    rel_index r = *this;
    r.off = v;
    r.rel = false;
    return r;
  }
};

// constrained templates, so that `i + 1` picks these (exact match on both operands) over the built-in operator
// reached through the conversion to long long
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_index>::type operator+(rel_index a, const I b) {
  a.off += static_cast<long long>(b);
  return a;
}
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_index>::type operator+(const I b, rel_index a) {
  a.off += static_cast<long long>(b);
  return a;
}
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_index>::type operator-(rel_index a, const I b) {
  a.off -= static_cast<long long>(b);
  return a;
}
// (centre + off) % m.  With m the lattice length and the cell's whole neighbourhood away from the periodic seam
// (the only place the tiled kernel uses rel_index), the result is the cell `off mod m` away, taken to the near side.
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_index>::type operator%(const rel_index a, const I m) {
  const long long mm = static_cast<long long>(m);
  if (a.rel && mm == static_cast<long long>(a.cells)) {
    rel_index r = a;
    long long o = a.off % mm;
    if (o > mm / 2) o -= mm;
    else if (o < -(mm / 2)) o += mm;
    r.off = o;
    return r;
  }
  return a.absolute(a.value() % mm);
}

template <class V> ODEINTR_DEV V loop_index(const long long c) { return static_cast<V>(c); }
template <class V> ODEINTR_DEV rel_index loop_index(const rel_index& c) { return c; }

}  // namespace odeintr_b200
