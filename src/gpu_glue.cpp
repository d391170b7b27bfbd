// src/gpu_glue.cpp -- the ONE fixed Rcpp translation unit a maintainer of thk686/odeintr adds so that the R closures
// NAME(), NAME_at(), ... of a compiled system call the B200 engine instead of a module built by Rcpp::sourceCpp.
// It replaces, call for call, what the reference generates from inst/templates/compile_sys_template.cpp:58-171 and
// the pars_*_template.cpp files: every export below names the generated function it stands for.  The engine is bound
// through the C-ABI of include/odeintr_b200.h (link: -L<repo>/odeintr_b200 -lodeintr_b200).
//
// R, Rcpp and a GPU are not available in the image this repository is built in, so this file is compiled only against
// the interface mock tests/mock_rcpp/Rcpp.h (tests/test_abi.py::test_rcpp_glue_compiles_against_the_header); the
// executed equivalent is odeintr_b200/_abi.py + system.py (ctypes), which every parity test goes through.
// [[Rcpp::plugins(cpp11)]]
#include <Rcpp.h>

#include <string>
#include <vector>

#include "odeintr_b200.h"

namespace {
void check(int rc) {
  if (rc != 0) Rcpp::stop(odeintr_last_error());
}
void finalize(odeintr_system_s* h) { odeintr_destroy(h); }
typedef Rcpp::XPtr<odeintr_system_s, Rcpp::PreserveStorage, finalize> handle_t;
}  // namespace

// replaces Rcpp::sourceCpp(code = code, env = env)                  R/odeintr.R:340, 457, 602
// [[Rcpp::export]]
SEXP odeintr_gpu_compile(std::string code, std::string name, int sys_dim, int n_pars, double atol, double rtol,
                         int flags) {
  odeintr_system_t h = nullptr;
  check(odeintr_compile(code.c_str(), name.c_str(), sys_dim, n_pars, atol, rtol, static_cast<unsigned>(flags), &h));
  return handle_t(h, true);
}

// NAME_get_output()                                                 compile_sys_template.cpp:58-73
// [[Rcpp::export]]
Rcpp::List odeintr_gpu_get_output(SEXP hp, int sys_dim) {
  handle_t h(hp);
  const size_t rows = odeintr_output_rows(h);
  Rcpp::NumericVector t(rows);
  check(odeintr_get_output_times(h, t.begin(), rows));
  std::vector<double> x(rows * sys_dim);
  check(odeintr_get_output(h, x.data(), x.size(), 0, 1));  // columns X1..XN, each `rows` long
  Rcpp::List out;
  out("Time") = t;
  for (int i = 0; i != sys_dim; ++i)
    out(std::string("X") + std::to_string(i + 1)) =
        Rcpp::NumericVector(x.begin() + i * rows, x.begin() + (i + 1) * rows);
  out.attr("class") = "data.frame";
  out.attr("row.names") = Rcpp::IntegerVector::create(NA_INTEGER, -static_cast<int>(rows));
  return out;
}

// NAME_set_state(new_state) / NAME_get_state()                      compile_sys_template.cpp:75-90
// [[Rcpp::export]]
void odeintr_gpu_set_state(SEXP hp, Rcpp::NumericVector x, int n_inst = 1) {
  check(odeintr_set_state(handle_t(hp), x.begin(), x.size(), n_inst));  // "Invalid initial state"
}
// [[Rcpp::export]]
std::vector<double> odeintr_gpu_get_state(SEXP hp, int sys_dim) {
  handle_t h(hp);
  std::vector<double> x(sys_dim * odeintr_n_instances(h));
  check(odeintr_get_state(h, x.data(), x.size()));
  return x;
}

// NAME_reset_observer()                                             compile_sys_template.cpp:92-97
// [[Rcpp::export]]
void odeintr_gpu_reset_observer(SEXP hp) { check(odeintr_reset_observer(handle_t(hp))); }

// NAME_set_params(a = dflt, ...) / NAME_get_params()                pars_setter_template.cpp:1-5, pars_getter_template.cpp:1-7
// (one value per parameter, or one per instance and parameter: p is n_pars x n_inst, instance index fastest)
// [[Rcpp::export]]
void odeintr_gpu_set_params(SEXP hp, Rcpp::NumericVector p, int n_inst = 1) {
  check(odeintr_set_params(handle_t(hp), p.begin(), p.size(), n_inst));  // "Invalid parameter vector"
}
// [[Rcpp::export]]
std::vector<double> odeintr_gpu_get_params(SEXP hp, int n_pars) {
  handle_t h(hp);
  std::vector<double> p(n_pars * odeintr_n_param_sets(h));
  check(odeintr_get_params(h, p.data(), p.size()));
  return p;
}

// NAME(init, duration, step_size = 1.0, start = 0.0)                compile_sys_template.cpp:141-158
// [[Rcpp::export]]
Rcpp::List odeintr_gpu_run(SEXP hp, Rcpp::NumericVector init, double duration, double step_size = 1.0,
                           double start = 0.0) {
  handle_t h(hp);
  if (duration / step_size < 3) {
    step_size = duration / 3;
    Rcpp::warning("Step size too large -- adjusting");
  }
  check(odeintr_set_state(h, init.begin(), init.size(), 1));
  check(odeintr_reset_observer(h));
  check(odeintr_integrate_const(h, start, start + duration, step_size, 1, 1));
  return odeintr_gpu_get_output(hp, init.size());
}

// NAME_adap(init, duration, step_size = 1.0, start = 0.0)           compile_sys_template.cpp:99-110
// [[Rcpp::export]]
Rcpp::List odeintr_gpu_adap(SEXP hp, Rcpp::NumericVector init, double duration, double step_size = 1.0,
                            double start = 0.0) {
  handle_t h(hp);
  check(odeintr_set_state(h, init.begin(), init.size(), 1));
  check(odeintr_reset_observer(h));
  check(odeintr_integrate_adaptive(h, start, start + duration, step_size, 1));
  const int n = init.size();
  if (!odeintr_output_is_ragged(h)) return odeintr_gpu_get_output(hp, n);
  // controlled / dense-output steppers: one row per accepted step, the Time column belongs to the trajectory
  const size_t cap = odeintr_output_rows(h);
  int rows = 0;
  check(odeintr_get_output_rows_per_instance(h, &rows, 1));
  std::vector<double> t(cap), x(cap * n);
  check(odeintr_get_output_times_per_instance(h, t.data(), cap, 0));
  check(odeintr_get_output(h, x.data(), x.size(), 0, 1));
  Rcpp::List out;
  out("Time") = Rcpp::NumericVector(t.begin(), t.begin() + rows);
  for (int i = 0; i != n; ++i)
    out(std::string("X") + std::to_string(i + 1)) =
        Rcpp::NumericVector(x.begin() + i * cap, x.begin() + i * cap + rows);
  out.attr("class") = "data.frame";
  out.attr("row.names") = Rcpp::IntegerVector::create(NA_INTEGER, -rows);
  return out;
}

// the tail NAME_at and NAME_continue_at share                       compile_sys_template.cpp:113-139
static Rcpp::List advance_and_observe(SEXP hp, int sys_dim, std::vector<double> times, double step_size, double start) {
  handle_t h(hp);
  if (times.empty()) Rcpp::stop("Invalid time specification");
  check(odeintr_integrate_const(h, start, times[0], step_size, 1, 0));  // whole steps to the first time, no observer
  check(odeintr_integrate_times(h, times.data(), times.size(), step_size));
  return odeintr_gpu_get_output(hp, sys_dim);
}

// NAME_at(init, times, step_size = 1.0, start = 0.0)                compile_sys_template.cpp:113-125
// [[Rcpp::export]]
Rcpp::List odeintr_gpu_at(SEXP hp, Rcpp::NumericVector init, std::vector<double> times, double step_size = 1.0,
                          double start = 0.0) {
  handle_t h(hp);
  check(odeintr_set_state(h, init.begin(), init.size(), 1));
  check(odeintr_reset_observer(h));
  return advance_and_observe(hp, init.size(), times, step_size, start);
}

// NAME_continue_at(times, step_size = 1.0)                          compile_sys_template.cpp:128-139
// [[Rcpp::export]]
Rcpp::List odeintr_gpu_continue_at(SEXP hp, int sys_dim, std::vector<double> times, double step_size = 1.0) {
  handle_t h(hp);
  const size_t rows = odeintr_output_rows(h);
  if (rows == 0) Rcpp::stop("no previous record to continue from");  // the reference reads rec_t.back() of an empty vector
  std::vector<double> t(rows);
  double start;
  if (odeintr_output_is_ragged(h)) {  // after NAME_adap: the last row of the trajectory's own Time column
    int n = 0;
    check(odeintr_get_output_rows_per_instance(h, &n, 1));
    check(odeintr_get_output_times_per_instance(h, t.data(), rows, 0));
    start = t[n - 1];
  } else {
    check(odeintr_get_output_times(h, t.data(), rows));
    start = t[rows - 1];
  }
  check(odeintr_reset_observer(h));
  return advance_and_observe(hp, sys_dim, times, step_size, start);
}

// NAME_no_record(init, duration, step_size = 1.0, start = 0.0)      compile_sys_template.cpp:160-171
// [[Rcpp::export]]
std::vector<double> odeintr_gpu_no_record(SEXP hp, Rcpp::NumericVector init, double duration, double step_size = 1.0,
                                          double start = 0.0) {
  handle_t h(hp);
  check(odeintr_set_state(h, init.begin(), init.size(), 1));
  check(odeintr_integrate_adaptive(h, start, start + duration, step_size, 0));
  std::vector<double> x(init.size());
  check(odeintr_get_state(h, x.data(), x.size()));  // the final state also stays in the engine for NAME_get_state()
  return x;
}

// compile_sys_openmp(..., halo = ): stencil radius of a large system (0 = stage-per-launch kernels, < 0 = measured)
// [[Rcpp::export]]
void odeintr_gpu_set_halo(SEXP hp, int halo, int fields = 1) { check(odeintr_lattice_set_halo(handle_t(hp), halo, fields)); }
