// tests/mock_rcpp/Rcpp.h -- TEST INFRASTRUCTURE ONLY.  The smallest stand-in for the part of Rcpp's interface that
// src/gpu_glue.cpp uses, so that the glue can be type-checked against include/odeintr_b200.h in an image without R
// (g++ -fsyntax-only).  Nothing here does anything.
#pragma once
#include <cstddef>
#include <string>
#include <vector>
typedef void* SEXP;
#define NA_INTEGER (-2147483647 - 1)
namespace Rcpp {
struct PreserveStorage {};
[[noreturn]] inline void stop(const std::string&) { throw 0; }
inline void warning(const std::string&) {}
struct NumericVector {
  std::vector<double> v;
  NumericVector() {}
  explicit NumericVector(std::size_t n) : v(n) {}
  template <class It> NumericVector(It a, It b) : v(a, b) {}
  double* begin() { return v.data(); }
  std::size_t size() const { return v.size(); }
};
struct IntegerVector {
  static IntegerVector create(int, int) { return IntegerVector(); }
};
struct List {
  struct Slot {
    template <class T> Slot& operator=(const T&) { return *this; }
  };
  Slot operator()(const std::string&) { return Slot(); }
  Slot attr(const std::string&) { return Slot(); }
};
template <class T, class Storage, void Finalizer(T*)>
struct XPtr {
  T* p;
  XPtr(T* p_, bool) : p(p_) {}
  XPtr(SEXP s) : p(static_cast<T*>(s)) {}
  operator T*() const { return p; }
  operator SEXP() const { return p; }
};
}  // namespace Rcpp
