// odeintr_b200/csrc/kernels/time_compare.cuh
// Boost's util/detail/less_with_sign.hpp (absolute epsilon), usable from host and device
// code.  Deliberately free of any other include so the host side of the library can use it too.
#pragma once

#define ODEINTR_HD __host__ __device__ __forceinline__
#define ODEINTR_EPS 2.220446049250313e-16

namespace odeintr_b200 {
ODEINTR_HD bool less_with_sign(double t1, double t2, double dt) {
  return dt > 0 ? (t2 - t1 > ODEINTR_EPS) : (t1 - t2 > ODEINTR_EPS);
}
ODEINTR_HD bool less_eq_with_sign(double t1, double t2, double dt) {
  return dt > 0 ? (t1 - t2 <= ODEINTR_EPS) : (t2 - t1 <= ODEINTR_EPS);
}
}  // namespace odeintr_b200
