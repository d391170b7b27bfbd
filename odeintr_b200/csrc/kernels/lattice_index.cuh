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
#define ODEINTR_HAVE_REL_INDEX 1  // lattice_helpers.cuh adds the overloads for symbolic coordinates

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
  bool wrapped = false;  // came out of `% cells`: as a NUMBER it is (centre + off) mod cells

  ODEINTR_DEV long long value() const {
    if (!rel) return off;
    long long v = static_cast<long long>(centre) + off;
    if (wrapped) {
      v %= cells;
      if (v < 0) v += cells;
    }
    return v;
  }
  ODEINTR_DEV operator long long() const { return value(); }
  ODEINTR_DEV rel_index absolute(const long long v) const {
    rel_index r = *this;
    r.off = v;
    r.rel = false;
    r.wrapped = false;
    return r;
  }
};

// constrained templates, so that `i + 1` picks these (exact match on both operands) over the built-in operator
// reached through the conversion to long long
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_index>::type operator+(rel_index a, const I b) {
  if (a.rel && a.wrapped) a = a.absolute(a.value());  // arithmetic after the wrap: a plain number again
  a.off += static_cast<long long>(b);
  return a;
}
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_index>::type operator+(const I b, rel_index a) {
  return a + b;
}
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_index>::type operator-(rel_index a, const I b) {
  if (a.rel && a.wrapped) a = a.absolute(a.value());
  a.off -= static_cast<long long>(b);
  return a;
}
// (centre + off) % m.  With m the lattice length and the cell's whole neighbourhood away from the periodic seam
// (the only place the tiled kernel uses rel_index), the result is the cell `off mod m` away, taken to the near side.
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_index>::type operator%(const rel_index a, const I m) {
  const long long mm = static_cast<long long>(m);
  if (a.rel && !a.wrapped && mm == static_cast<long long>(a.cells)) {
    rel_index r = a;
    long long o = a.off % mm;
    if (o > mm / 2) o -= mm;
    else if (o < -(mm / 2)) o += mm;
    r.off = o;
    r.wrapped = true;
    return r;
  }
  return a.absolute(a.value() % mm);
}

template <class V> ODEINTR_DEV V loop_index(const long long c) { return static_cast<V>(c); }
template <class V> ODEINTR_DEV rel_index loop_index(const rel_index& c) { return c; }

}  // namespace odeintr_b200

// ---- M x M lattices (laplace2D & friends, inst/include/utils.h:6-31) ------------------------------------------------
// The reference's large-system example walks a flat index i over an M x M row-major lattice and recovers the
// coordinates as `i / M`, `i % M`; its helpers wrap neighbours with bound() and index x[row * M + col].  The same
// symbolic treatment as above, per coordinate: rel_coord is "the cell's own row (column) + displacement", bound()
// keeps the displacement (the nearest periodic image), `row * M + col` assembles a rel_cell2 that the 2-D tile view
// serves as one shared-memory load at a fixed (d_row, d_col) offset.  Everything else degrades to absolute values.
namespace odeintr_b200 {

struct rel_coord {
  long long off;   // rel: displacement from the cell's own coordinate; otherwise the absolute value
  int centre;      // the cell's own coordinate
  int side;        // M
  bool rel;
  bool row;        // a row number (true) or a column number (false)
  bool wrapped;    // went through bound(): the absolute value is taken modulo M
  ODEINTR_DEV long long value() const {
    if (!rel) return off;
    long long v = static_cast<long long>(centre) + off;
    if (wrapped) {
      v %= side;
      if (v < 0) v += side;
    }
    return v;
  }
  ODEINTR_DEV operator long long() const { return value(); }
};

template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_coord>::type operator+(rel_coord a, const I b) {
  if (a.rel && a.wrapped) { a.off = a.value(); a.rel = false; }  // arithmetic after bound(): plain numbers again
  a.off += static_cast<long long>(b);
  return a;
}
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_coord>::type operator+(const I b, const rel_coord a) {
  return a + b;
}
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_coord>::type operator-(const rel_coord a, const I b) {
  return a + (-static_cast<long long>(b));
}

// flat index (row * M + col) being assembled from symbolic coordinates
struct rel_cell2 {
  long long dr, dc;        // displacements (rel)
  long long abs;           // absolute flat value of the parts seen so far
  bool rel, has_row, has_col;
  ODEINTR_DEV operator long long() const { return abs; }
};

// row * M
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_cell2>::type operator*(const rel_coord a, const I m) {
  rel_cell2 c;
  c.abs = a.value() * static_cast<long long>(m);
  c.rel = a.rel && a.row && static_cast<long long>(m) == a.side;
  c.dr = a.off; c.dc = 0;
  c.has_row = true; c.has_col = false;
  return c;
}
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_cell2>::type operator*(const I m, const rel_coord a) {
  return a * m;
}
// row * M + col
ODEINTR_DEV rel_cell2 operator+(rel_cell2 c, const rel_coord b) {
  c.abs += b.value();
  c.rel = c.rel && b.rel && !b.row && c.has_row && !c.has_col;
  c.dc = b.off;
  c.has_col = true;
  return c;
}
ODEINTR_DEV rel_cell2 operator+(const rel_coord b, const rel_cell2 c) { return c + b; }
// anything else added to a flat index: a plain number
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_cell2>::type operator+(rel_cell2 c, const I b) {
  c.abs += static_cast<long long>(b);
  c.rel = false;
  return c;
}

// The loop variable on an M x M lattice: the flat cell number with its coordinates.  `i / M` and `i % M` give the
// symbolic row and column; any other arithmetic sees the plain flat number.
struct rel_index2 {
  int centre;  // flat number of the cell being computed (row * M + col)
  int row, col;
  int side;    // M
  ODEINTR_DEV operator long long() const { return centre; }
};
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_coord>::type operator/(const rel_index2& i, const I m) {
  rel_coord r;
  r.side = i.side; r.row = true; r.wrapped = false;
  r.rel = static_cast<long long>(m) == i.side;
  r.centre = i.row;
  r.off = r.rel ? 0 : static_cast<long long>(i.centre) / static_cast<long long>(m);
  return r;
}
template <class I>
ODEINTR_DEV typename enable_if_<is_integer<I>::value, rel_coord>::type operator%(const rel_index2& i, const I m) {
  rel_coord r;
  r.side = i.side; r.row = false; r.wrapped = false;
  r.rel = static_cast<long long>(m) == i.side;
  r.centre = i.col;
  r.off = r.rel ? 0 : static_cast<long long>(i.centre) % static_cast<long long>(m);
  return r;
}
template <class V> ODEINTR_DEV rel_index2 loop_index(const rel_index2& c) { return c; }

}  // namespace odeintr_b200

// std::min / std::max of a symbolic index and a plain integer (clamped boundaries such as x[std::max(i - 1, 0)]):
// the function templates of <algorithm> need both arguments of one type, which held while the loop variable was an
// int.  The result is an ordinary number; the access takes the views' absolute-index path.
namespace std {
#define ODEINTR_MIXED_MINMAX_(TYPE)                                                                              \
  template <class I>                                                                                             \
  ODEINTR_DEV typename odeintr_b200::enable_if_<odeintr_b200::is_integer<I>::value, long long>::type min(        \
      const TYPE& a, const I b) { const long long v = a; return static_cast<long long>(b) < v ? static_cast<long long>(b) : v; } \
  template <class I>                                                                                             \
  ODEINTR_DEV typename odeintr_b200::enable_if_<odeintr_b200::is_integer<I>::value, long long>::type min(        \
      const I b, const TYPE& a) { const long long v = a; return v < static_cast<long long>(b) ? v : static_cast<long long>(b); } \
  template <class I>                                                                                             \
  ODEINTR_DEV typename odeintr_b200::enable_if_<odeintr_b200::is_integer<I>::value, long long>::type max(        \
      const TYPE& a, const I b) { const long long v = a; return v < static_cast<long long>(b) ? static_cast<long long>(b) : v; } \
  template <class I>                                                                                             \
  ODEINTR_DEV typename odeintr_b200::enable_if_<odeintr_b200::is_integer<I>::value, long long>::type max(        \
      const I b, const TYPE& a) { const long long v = a; return static_cast<long long>(b) < v ? v : static_cast<long long>(b); }
ODEINTR_MIXED_MINMAX_(odeintr_b200::rel_index)
ODEINTR_MIXED_MINMAX_(odeintr_b200::rel_index2)
ODEINTR_MIXED_MINMAX_(odeintr_b200::rel_coord)
#undef ODEINTR_MIXED_MINMAX_
}  // namespace std
