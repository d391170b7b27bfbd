// odeintr_b200/csrc/builtin_kernels.h — host launchers of the statically compiled helper kernels.
#pragma once
#include <cuda_runtime.h>

namespace odeintr_b200 {
void launch_fill_state_uniform(double* x, long long m, long long ld, int n, unsigned long long seed,
                               unsigned long long first, const double* lohi_dev, cudaStream_t st);
void launch_fill_params_linspace(double* p, long long m, long long ld, int n, unsigned long long first,
                                 unsigned long long total, const double* lohi_dev, cudaStream_t st);
void launch_gather_output(const double* out, double* dst, long long rows, int n, long long ld, long long inst_begin,
                          long long inst_count, cudaStream_t st);
void launch_broadcast_state(double* x, long long m, long long ld, int n, const double* src_dev, cudaStream_t st);
void launch_fp64_probe(int kind, double* sink, int iters, unsigned blocks, cudaStream_t st);
void launch_copy_probe(const void* src, void* dst, long long bytes, unsigned blocks, cudaStream_t st);
// K6 instance order: stable radix sort of (key, instance index) pairs (CUB); vals_in = 0, 1, 2, ... is filled here
size_t sort_pairs_temp_bytes(long long n);
int launch_sort_by_key(void* temp, size_t temp_bytes, const unsigned* keys_in, unsigned* keys_out, int* vals_in,
                       int* vals_out, long long n, cudaStream_t st);
void launch_exact_div_check(unsigned long long pairs_per_thread, unsigned blocks, unsigned long long seed,
                            unsigned long long* mismatches_dev, cudaStream_t st);
void launch_fill_negzero(double* v, long long n, cudaStream_t st);
void philox_host(unsigned long long counter, unsigned block, unsigned long long seed, unsigned out[4]);
}  // namespace odeintr_b200
