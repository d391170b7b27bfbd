// odeintr_b200/csrc/kernels/integrate_device.cuh
// Per-trajectory use of the integrate loops (integrate_loops.cuh): every thread runs the loop of its own
// trajectory, so step sizes, rejected tries and the number of real steps between two observations are per
// instance.  This header adds the device observer.
#pragma once
#include "odeintr_b200/adaptive_steppers.cuh"
#include "odeintr_b200/rk78.cuh"
#include "odeintr_b200/bulirsch_stoer.cuh"
#include "odeintr_b200/integrate_loops.cuh"

namespace odeintr_b200 {

// Observer: one SoA row per variable per sample, instance index fastest (coalesced across a warp whenever
// neighbouring trajectories reach the same sample together, which they do for smooth sweeps).
template <int N>
struct sample_writer {
  double* out;        // points at this instance's column of sample 0 (or null: null_observer)
  double* t_rec;      // per-instance sample times (ragged adaptive records) or null
  long long ld;
  int rows, max_rows;
  ODEINTR_DEV void operator()(const vec<N>& x, const double t) {
    if (out && rows < max_rows) {
#pragma unroll
      for (int v = 0; v < N; ++v) store_stream(out + ((long long)rows * N + v) * ld, x[v]);
      if (t_rec) store_stream(t_rec + (long long)rows * ld, t);
    }
    ++rows;
  }
};

}  // namespace odeintr_b200
