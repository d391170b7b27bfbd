// odeintr_b200/csrc/kernels/exact_div.cuh
// Correctly rounded division by a divisor that is used many times.
//
// The dense output of dopri5 divides by six constants per sample (Boost's interpolation coefficients).  An IEEE double
// division is a reciprocal approximation, two Newton steps and a correction -- about ten FP64-pipe instructions plus a
// range check -- and most of that work depends on the divisor alone.  With the divisor's correctly rounded reciprocal
// y = RN(1 / b) at hand (a compile-time constant here), a / b follows from two correction steps:
//     q0 = RN(a y)                       |q0 - a/b| < 1.5 ulp
//     r0 = a - b q0   (exact, one FMA)   q1 = RN(q0 + r0 y)      within 1/2 ulp + 2^-105 of a/b: a faithful rounding
//     r1 = a - b q1   (exact, one FMA)   q  = RN(q1 + r1 y)      = RN(a / b)
// The last line is Markstein's theorem (P. Markstein, IBM J. Res. Dev. 34, 1990; Muller et al., Handbook of
// Floating-Point Arithmetic, sec. 4.7): for y = RN(1/b) and a faithful q1 the corrected quotient is the correctly
// rounded one -- the same bits as a / b, hence as the x86 divsd of the reference's host code.  The residuals are exact
// only while nothing under- or overflows on the way, so operands outside a wide exponent window (and zeros,
// infinities, NaNs) take the ordinary division.  `fma` is explicit here; --fmad=false only forbids contracting a * b + c
// written as separate operations.  odeintr_selftest_exact_division compares 2^28 quotients with a / b on the device.
// Measured (round 2): +1.7 % on config 3 (one sample per ~4 attempts).  For run-time divisors (rosenbrock4's step size
// and LU pivots: one true division for y, then 5 operations per quotient) it was NOT faster than the device's own
// division sequence and cost registers, so rosenbrock4 keeps plain divisions.
#pragma once

namespace odeintr_b200 {

// 2^-lim <= |v| <= 2^lim, from the exponent field alone (false for zeros, denormals, infinities, NaNs)
__device__ __forceinline__ bool exponent_within(const double v, const int lim) {
  const unsigned e = (static_cast<unsigned>(__double2hiint(v)) >> 20) & 0x7ffu;
  return e - static_cast<unsigned>(1023 - lim) <= static_cast<unsigned>(2 * lim);
}

// the rare operands outside the window: one out-of-line IEEE division for all call sites
__device__ __noinline__ double ieee_division(const double a, const double b) { return a / b; }

struct divisor {
  double b, y;
  bool tame;
  __device__ __forceinline__ divisor() : b(1.0), y(1.0), tame(true) {}
  __device__ __forceinline__ explicit divisor(const double b_) : b(b_), y(1.0 / b_), tame(exponent_within(b_, 480)) {}
  // compile-time divisors: y = RN(1 / b) folded by the compiler
  __device__ __forceinline__ constexpr divisor(const double b_, const double y_) : b(b_), y(y_), tame(true) {}
  // a / b
  __device__ __forceinline__ double under(const double a) const {
    if (tame && exponent_within(a, 480)) {
      const double q0 = a * y;
      const double q1 = fma(fma(-b, q0, a), y, q0);
      return fma(fma(-b, q1, a), y, q1);
    }
    return ieee_division(a, b);
  }
};

}  // namespace odeintr_b200
