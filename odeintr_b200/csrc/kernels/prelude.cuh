// odeintr_b200/csrc/kernels/prelude.cuh
// Device-side prelude seen by every generated translation unit before the user's snippet.
// It supplies the names a snippet written for the reference's host templates can use
// (/root/reference inst/templates/compile_sys_template.cpp:15-39): std:: math, std::size_t,
// a fixed-size state_type with operator[] and size(), and min/max.  NVRTC has no host headers,
// so nothing here includes one.
#pragma once
#include "odeintr_b200/abi_args.h"
#include "odeintr_b200/time_compare.cuh"
#include "odeintr_b200/exact_div.cuh"

#ifndef ODEINTR_BLOCK
#define ODEINTR_BLOCK 128
#endif
#ifndef ODEINTR_MIN_BLOCKS   // resident blocks per SM the ensemble kernels are compiled for (register cap)
#define ODEINTR_MIN_BLOCKS 1
#endif

#define ODEINTR_DEV __device__ __forceinline__

namespace std {
typedef unsigned long size_t;
typedef long ptrdiff_t;
using ::sin; using ::cos; using ::tan; using ::asin; using ::acos; using ::atan; using ::atan2;
using ::sinh; using ::cosh; using ::tanh; using ::asinh; using ::acosh; using ::atanh;
using ::exp; using ::exp2; using ::expm1; using ::log; using ::log2; using ::log10; using ::log1p;
using ::sqrt; using ::cbrt; using ::pow; using ::hypot; using ::fabs; using ::floor; using ::ceil;
using ::trunc; using ::round; using ::fmod; using ::fmin; using ::fmax; using ::erf; using ::erfc;
using ::tgamma; using ::lgamma; using ::copysign; using ::isnan; using ::isinf; using ::isfinite;
ODEINTR_DEV double abs(double a) { return ::fabs(a); }
ODEINTR_DEV int abs(int a) { return a < 0 ? -a : a; }
template <class T> ODEINTR_DEV const T& min(const T& a, const T& b) { return b < a ? b : a; }
template <class T> ODEINTR_DEV const T& max(const T& a, const T& b) { return a < b ? b : a; }
}  // namespace std

namespace odeintr_b200 {

// Register-resident state vector.  All indexing in the steppers is by compile-time constants after
// unrolling, so `v` is promoted to registers (one trajectory per thread).
template <int N_>
struct vec {
  double v[N_];
  ODEINTR_DEV double& operator[](int i) { return v[i]; }
  ODEINTR_DEV const double& operator[](int i) const { return v[i]; }
  ODEINTR_DEV constexpr std::size_t size() const { return N_; }
};

// N x N matrix for the implicit path; J(i, j) syntax as in uBLAS
// (/root/reference inst/templates/compile_implicit_template.cpp:49-56).
template <int N_>
struct mat {
  double a[N_][N_];
  ODEINTR_DEV double& operator()(int i, int j) { return a[i][j]; }
  ODEINTR_DEV const double& operator()(int i, int j) const { return a[i][j]; }
};

constexpr int isqrt(unsigned long n) {
  unsigned long r = 0;
  while ((r + 1) * (r + 1) <= n) ++r;
  return static_cast<int>(r);
}

// streaming (evict-first) stores for observer samples: they are written once and never re-read
// by the kernel, so they should not displace anything in L2.
ODEINTR_DEV void store_stream(double* p, double v) { __stcs(p, v); }

}  // namespace odeintr_b200
