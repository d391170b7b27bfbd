// odeintr_b200/csrc/abi.cu — implementation of the C-ABI declared in include/odeintr_b200.h.
//
// Host side of the hot path: NVRTC-compiles the generated translation unit (user __device__ functor +
// hand-written stepper kernels from csrc/kernels/, embedded in this library as strings), owns the
// device-resident ensemble (state, parameters, observer record) and turns the reference's
// integrate_const / integrate_times / integrate_adaptive calls
// (/root/reference inst/templates/compile_sys_template.cpp:107,121,123,134,136,154,168) into kernel
// launches.  Step schedules that do not depend on the state (fixed-step steppers) are computed here
// with Boost's own loop conditions so the device executes exactly the reference's step sequence.
//
// There is deliberately no CPU path: without a device every compute entry point fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvrtc.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/odeintr_b200.h"
#include "builtin_kernels.h"
#include "kernels/abi_args.h"
#include "kernels/integrate_loops.cuh"
#include "embedded_headers.inc"  // generated: k_header_names[], k_header_srcs[], k_num_headers

namespace {

thread_local std::string g_err;
std::atomic<long long> g_total_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// NCCL is bound at run time (dlopen) so that the library also loads where NCCL is absent; inside a process that
// already imported torch this resolves to the libnccl torch loaded.  Only the halo exchange of sharded lattices
// uses it (ncclSend / ncclRecv to the two ring neighbours inside one group).
struct nccl_api {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
nccl_api g_nccl;

int load_nccl() {
  if (g_nccl.ok) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return fail(3, std::string("cannot load libnccl.so.2: ") + dlerror());
#define ODEINTR_NCCL_SYM(field, name)                                                   \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                                   \
  if (!g_nccl.field) return fail(3, std::string("libnccl is missing ") + name);
  ODEINTR_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  ODEINTR_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  ODEINTR_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  ODEINTR_NCCL_SYM(GroupStart, "ncclGroupStart")
  ODEINTR_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  ODEINTR_NCCL_SYM(Send, "ncclSend")
  ODEINTR_NCCL_SYM(Recv, "ncclRecv")
  ODEINTR_NCCL_SYM(AllReduce, "ncclAllReduce")
  ODEINTR_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ODEINTR_NCCL_SYM
  g_nccl.ok = true;
  return 0;
}
#define NCCL_TRY(expr)                                                                          \
  do {                                                                                          \
    ncclResult_t r_ = (expr);                                                                   \
    if (r_ != ncclSuccess) return fail(ODEINTR_E_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(r_)); \
  } while (0)
#define CU_TRY(expr)                                                                        \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(ODEINTR_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));      \
  } while (0)

// ---- util/detail/less_with_sign.hpp (Boost), SURVEY.md Appendix A --------------------------------
const double kEps = std::numeric_limits<double>::epsilon();
inline bool less_with_sign(double t1, double t2, double dt) { return dt > 0 ? (t2 - t1 > kEps) : (t1 - t2 > kEps); }
inline bool less_eq_with_sign(double t1, double t2, double dt) {
  return dt > 0 ? (t1 - t2 <= kEps) : (t2 - t1 <= kEps);
}
inline double min_abs(double a, double b) { return a > 0 ? std::min(a, b) : std::max(a, b); }

// number of steps integrate_const's plain loop takes: first `step` for which
// less_eq_with_sign(fl(t0 + step*dt) + dt, t1, dt) is false (the condition is monotone in step).
long long const_step_count(double t0, double t1, double dt) {
  auto cond = [&](long long s) { return less_eq_with_sign((t0 + static_cast<double>(s) * dt) + dt, t1, dt); };
  const double est = std::floor((t1 - t0) / dt);
  if (!(est >= 0)) return 0;
  long long s = est > 4 ? static_cast<long long>(est) - 4 : 0;
  while (s > 0 && !cond(s - 1)) --s;
  while (cond(s)) ++s;
  return s;
}

template <class T>
struct dev_buf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// cells per thread and stage of the tiled rk4 kernel (ODEINTR_TILE_CPT in lattice_tiled.cuh; env knob for sweeps)
int tile_cpt() {
  const char* c = std::getenv("ODEINTR_TILE_CPT");
  const int v = c ? std::atoi(c) : 4;
  return v > 0 ? v : 4;
}

// rows of a tile of the 2-D tiled rk4 kernel (ODEINTR_TILE2_ROWS in lattice_tiled2d.cuh; env knob for sweeps)
int tile2_rows() {
  const char* c = std::getenv("ODEINTR_TILE2_ROWS");
  const int v = c ? std::atoi(c) : 32;
  return (v > 0 && v % 4 == 0) ? v : 32;
}

std::vector<std::string> nvrtc_options(unsigned flags) {
  std::vector<std::string> o;
  o.push_back("--gpu-architecture=sm_100a");
  o.push_back("--std=c++17");
  o.push_back("-default-device");
  o.push_back("-lineinfo");
  o.push_back((flags & ODEINTR_FLAG_FMAD) ? "--fmad=true" : "--fmad=false");
  if (flags & ODEINTR_FLAG_KEEP_ZERO_TERMS) o.push_back("-DODEINTR_KEEP_ZERO_TERMS=1");
  if (flags & ODEINTR_FLAG_RECIPROCAL) o.push_back("-DODEINTR_RECIPROCAL=1");
  // tuning knobs for experiments (scripts/explore_*.py): register cap and block size of the ensemble kernels
  if (const char* r = std::getenv("ODEINTR_MAXRREGCOUNT")) o.push_back(std::string("--maxrregcount=") + r);
  if (const char* b = std::getenv("ODEINTR_BLOCK")) o.push_back(std::string("-DODEINTR_BLOCK=") + b);
  if (const char* b = std::getenv("ODEINTR_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_MIN_BLOCKS=") + b);
  if (std::getenv("ODEINTR_NESTED_LOOPS")) o.push_back("-DODEINTR_NESTED_LOOPS=1");
  if (const char* c = std::getenv("ODEINTR_TILE_CPT")) o.push_back(std::string("-DODEINTR_TILE_CPT=") + c);
  if (const char* c = std::getenv("ODEINTR_TILE_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_TILE_MIN_BLOCKS=") + c);
  if (const char* c = std::getenv("ODEINTR_TMA_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_TMA_MIN_BLOCKS=") + c);
  if (const char* c = std::getenv("ODEINTR_WARP5_CPT")) o.push_back(std::string("-DODEINTR_WARP5_CPT=") + c);
  if (const char* c = std::getenv("ODEINTR_WARP5_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_WARP5_MIN_BLOCKS=") + c);
  if (const char* c = std::getenv("ODEINTR_STREAM2_L2_AHEAD")) o.push_back(std::string("-DODEINTR_STREAM2_L2_AHEAD=") + c);
  if (std::getenv("ODEINTR_STREAM5_LATE_NORM")) o.push_back("-DODEINTR_STREAM5_LATE_NORM=1");
  if (std::getenv("ODEINTR_DP5_LITERALS")) o.push_back("-DODEINTR_DP5_LITERALS=1");
  if (const char* c = std::getenv("ODEINTR_ENS_ADAPTIVE_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_ENS_ADAPTIVE_MIN_BLOCKS=") + c);
  if (const char* c = std::getenv("ODEINTR_STREAM5_THREADS")) o.push_back(std::string("-DODEINTR_STREAM5_THREADS=") + c);
  if (const char* c = std::getenv("ODEINTR_STREAM5_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_STREAM5_MIN_BLOCKS=") + c);
  if (const char* c = std::getenv("ODEINTR_STREAM2_ROWS")) o.push_back(std::string("-DODEINTR_STREAM2_ROWS=") + c);
  if (const char* c = std::getenv("ODEINTR_STREAM2_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_STREAM2_MIN_BLOCKS=") + c);
  if (const char* c = std::getenv("ODEINTR_WARP_CPT")) o.push_back(std::string("-DODEINTR_WARP_CPT=") + c);
  if (const char* c = std::getenv("ODEINTR_WARP_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_WARP_MIN_BLOCKS=") + c);
  if (const char* c = std::getenv("ODEINTR_TILE2_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_TILE2_MIN_BLOCKS=") + c);
  if (const char* c = std::getenv("ODEINTR_TILE5_CPT")) o.push_back(std::string("-DODEINTR_TILE5_CPT=") + c);
  if (const char* c = std::getenv("ODEINTR_TILE5_MIN_BLOCKS")) o.push_back(std::string("-DODEINTR_TILE5_MIN_BLOCKS=") + c);
  if (std::getenv("ODEINTR_TILE2_ROWS")) o.push_back("-DODEINTR_TILE2_ROWS=" + std::to_string(tile2_rows()));
  return o;
}

int nvrtc_to_cubin_once(const char* src, unsigned flags, bool staged_only, std::vector<char>& cubin, std::string& log);

// Large-system translation units are first built with every kernel family; if the snippet does not compile with
// the symbolic loop index the tiled kernels hand it (lattice_index.cuh), the unit is rebuilt with the
// stage-per-launch kernels alone, where the loop variable is the plain integer the reference's code sees.
int nvrtc_to_cubin(const char* src, unsigned flags, std::vector<char>& cubin, std::string& log) {
  int rc = nvrtc_to_cubin_once(src, flags, false, cubin, log);
  if (rc == ODEINTR_OK || !std::strstr(src, "ODEINTR_LATTICE_STAGES")) return rc;
  const std::string first_error = odeintr_last_error();
  std::string log2;
  if (nvrtc_to_cubin_once(src, flags, true, cubin, log2) == ODEINTR_OK) {
    log = "note: built with the stage-per-launch kernels only (the snippet does not compile with the symbolic loop "
          "index of the tiled kernels)\n" + log2;
    return ODEINTR_OK;
  }
  return fail(ODEINTR_E_COMPILE, first_error);  // the snippet itself is at fault: report the first diagnosis
}

int nvrtc_to_cubin_once(const char* src, unsigned flags, bool staged_only, std::vector<char>& cubin, std::string& log) {
  nvrtcProgram prog;
  nvrtcResult r = nvrtcCreateProgram(&prog, src, "odeintr_system.cu", k_num_headers, k_header_srcs, k_header_names);
  if (r != NVRTC_SUCCESS) return fail(ODEINTR_E_COMPILE, std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(r));
  std::vector<std::string> opts = nvrtc_options(flags);
  if (staged_only) opts.push_back("-DODEINTR_NO_SYMBOLIC=1");
  std::vector<const char*> copts;
  for (auto& s : opts) copts.push_back(s.c_str());
  r = nvrtcCompileProgram(prog, (int)copts.size(), copts.data());
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  log.assign(log_size, '\0');
  if (log_size) nvrtcGetProgramLog(prog, &log[0]);
  if (r != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    return fail(ODEINTR_E_COMPILE, std::string("NVRTC compilation failed: ") + nvrtcGetErrorString(r) + "\n" + log);
  }
  size_t sz = 0;
  r = nvrtcGetCUBINSize(prog, &sz);
  if (r != NVRTC_SUCCESS || sz == 0) {
    nvrtcDestroyProgram(&prog);
    return fail(ODEINTR_E_COMPILE, "nvrtcGetCUBINSize failed");
  }
  cubin.resize(sz);
  nvrtcGetCUBIN(prog, cubin.data());
  nvrtcDestroyProgram(&prog);
  return ODEINTR_OK;
}

}  // namespace

struct odeintr_system_s {
  std::string name, log;
  int N = 0, P = 0, device = 0;
  double atol = 1e-6, rtol = 1e-6;
  unsigned flags = 0;
  cudaLibrary_t lib = nullptr;
  cudaKernel_t k_plain = nullptr, k_adaptive = nullptr, k_lattice = nullptr, k_lattice_sharded = nullptr,
               k_combo = nullptr, k_tiled = nullptr, k_tiled_int = nullptr, k_tiled_tma = nullptr,
               k_info = nullptr, k_probe = nullptr, k_lat_ens = nullptr, k_tiled2d = nullptr, k_tiled2d_h[3] = {nullptr, nullptr, nullptr},
               k_dopri5_tiled = nullptr, k_dopri5_tiled_h[3] = {nullptr, nullptr, nullptr};
  cudaKernel_t k_warp_h[3] = {nullptr, nullptr, nullptr}, k_warp = nullptr, k_warp_info = nullptr;
  int warp_cpt = 0;             // cells per lane the warp kernels were compiled with (ODEINTR_WARP_CPT)
  int warp_halo = 0;            // radius of the selected warp kernel (1 or 2)
  size_t warp_smem = 0;         // dynamic shared memory of the warp kernel
  int warp_blocks = 0;          // its resident blocks per SM (0 = not available for the current radius)
  cudaStream_t stream2 = nullptr;        // sharded runs: the halo exchange travels here, beside the interior kernels
  cudaEvent_t ev_edges = nullptr, ev_xchg = nullptr;
  cudaEvent_t ev_mid = nullptr, ev_side = nullptr;  // fused step pairs: the middle (main stream) and the edges (second stream) of a pair run side by side
  cudaKernel_t k_pilot = nullptr;
  cudaKernel_t k_serial = nullptr;  // general large-system snippet: the right-hand side is one serial kernel (correctness path)
  int order_mode = 0;           // K6: 0 = instances in index order, 1 = sorted by the pilot's cost proxy, 2 = caller's permutation
  int pilot_attempts = 12;
  dev_buf<int> perm, perm_iota;
  dev_buf<unsigned> order_keys, order_keys_sorted;
  dev_buf<unsigned char> sort_temp;
  size_t perm_m = 0;            // instances the resident permutation was made for (mode 2)
  size_t lat_vec_len = 0;       // doubles per state-like vector of the adaptive large-system loops (N, or a slab)
  std::vector<const double*> slab_fresh;  // slab vectors whose ghost cells hold the neighbours' current values
  dev_buf<double> lat_negzero;  // -0.0 everywhere: (-0.0) + k == k bit for bit (RHS alone through the stage kernel)
  bool tiled_on = false;        // rk4 steps of this system go through the temporally blocked kernels
  bool tiled_auto = false;      // the radius was measured by the probe (a violation falls back to the staged kernels)
  bool tiled_pending = false;   // measure the radius at the next integrate call
  bool tiled2d_on = false;      // M x M lattice: rk4 steps go through odeintr_lattice_rk4_tiled2d
  cudaKernel_t k_stream2d = nullptr, k_stream2d_info = nullptr;  // M x M lattices, radius 1: streaming rk4 step (lattice_stream2d.cuh)
  int stream2d_rows = 0, stream2d_blocks = 0;
  int tiled5 = -1;              // dopri5 attempts go through odeintr_lattice_dopri5_tiled: -1 undecided, 0 no, 1 yes
  cudaKernel_t k_warp5_h[3] = {nullptr, nullptr, nullptr}, k_warp5 = nullptr, k_warp5_info = nullptr;
  int warp5_cpt = 0, warp5_halo = 0, warp5_blocks = 0;  // warp-autonomous dopri5 attempt (lattice_dopri5_warp.cuh)
  cudaKernel_t k_lat_ens_plain = nullptr;  // ensembles of wider systems under euler / rk54 / rk5 / rk78 (fixed step)
  cudaKernel_t k_lat_ens_a[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // adaptive ensembles of wider systems (lattice_ensemble_adaptive.cuh), by lattice_method
  cudaKernel_t k_stream5 = nullptr, k_stream5_info = nullptr;  // M x M lattices, radius 1: streaming dopri5 attempt (lattice_dopri5_stream2d.cuh); tiled5 == 2
  int stream5_blocks = 0, stream5_rows = 0, stream5_threads = 256;
  size_t warp5_smem = 0;
  long long tiled5_halo = 0;
  size_t tiled5_smem = 0;
  long long lat_side = 0;       // M when the snippet addresses an M x M lattice (one field), else 0
  size_t tiled2d_smem = 0;
  long long lattice_halo = 0;   // stencil radius the tiled kernels stage (declared, or measured by the probe)
  long long lattice_cells = 0;  // cells per field (sys_dim / fields) for the tiled kernel
  size_t tiled_smem = 0;        // dynamic shared memory of the general tiled kernel (4 arrays per field)
  size_t tiled_int_smem = 0;    // ... of the interior kernel (2 arrays per field)
  size_t tiled_tma_smem = 0;    // ... of the persistent TMA kernel (4 arrays per field + 2 mbarriers)
  int tiled_tma_blocks = 0;     // resident blocks per SM of the persistent TMA kernel (0 = not available)
  long long lat_start = 0, lat_bound = 0;  // range of the user's cell loop (odeintr_lattice_info)
  long long lat_cells_ct = 0;   // cells per field compiled into the interior kernel
  dev_buf<int> lat_flag;
  int lattice_method = 0;  // 0 = fixed step (rk4 / euler / rk54), 1 = rk5_a, 2 = rk5_i, 3 = rk54_a, 4 = rk78_a, 5 = bs, 6 = bsd
  dev_buf<double> lat_work;
  std::vector<double*> lat_pool;  // work vectors of the Bulirsch-Stoer steppers (their number depends on the orders reached)
  size_t lat_pool_len = 0, lat_pool_used = 0;
  bool lat_mid_primed = false;   // the intermediate state of fused step pairs holds x on the cells no loop writes
  bool lat_work_primed = false;  // the work vectors of the fixed-step rk54 / rk78 / rk5 paths hold copies of x / zeros
  int lat_rk5_parity = 0;        // which of the two derivative vectors holds k1 of the next rk5 step
  dev_buf<unsigned long long> lat_err;
  int block_plain = 128, block_adaptive = 128;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;

  dev_buf<double> x, pars, out, stage, t_rec;
  dev_buf<long long> seg_nfull, accepted, rejected;
  dev_buf<double> seg_hrem, seg_t, times_dev;
  dev_buf<unsigned char> seg_obs;
  dev_buf<int> status, rows_inst, fail_flags, rows_min;
  // large-system (lattice) path
  int lattice_stages = 4, sm_count = 148;
  dev_buf<double> lat_buf_a, lat_buf_b, lat_buf_acc;
  double *lat_x = nullptr, *lat_a = nullptr, *lat_b = nullptr, *lat_acc = nullptr;
  ncclComm_t comm = nullptr;
  int comm_world = 1, comm_rank = 0;
  bool sharded = false, sh_primed = false;
  long long sh_total = 0, sh_c0 = 0, sh_n = 0, sh_ghost = 0, sh_pitch = 0, sh_pad = 0;
  int sh_fields = 1, sh_halo = 1;
  int sh_interval = 1;          // rk4 steps per halo exchange requested for the tiled kernels (ghost = 4 * halo * interval)
  int sh_steps_per_exchange = 1;  // ... in effect for the current slab
  bool sh_tiled = false;        // the slab steps through the tiled kernels
  double* sh_x = nullptr;
  long long max_tries = 20000000;  // per-instance step-attempt budget per call
  bool dense = false;    // the adaptive kernel wraps a dense-output stepper (rk5_i, rosenbrock4)
  dev_buf<unsigned long long> work_counter;
  size_t m = 0;          // resident instances
  size_t par_sets = 0;   // 0 (none set), 1 (broadcast) or m
  std::vector<double> rec_t;  // Time column of the record
  size_t rows = 0;
  size_t out_m = 0;      // instance count the record was written with
  bool ragged = false;   // per-instance row counts (adaptive record)
  int max_rows = 0;
  long long last_steps = 0, last_launches = 0;
  // host-streaming pipeline (odeintr_integrate_*_host)
  bool pipe_ready = false;
  cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t pipe_done[3] = {nullptr, nullptr, nullptr};
  dev_buf<double> pipe_x[3], pipe_out[3], pipe_pars[3];
  dev_buf<double> pipe_bcast;  // one broadcast parameter set of a host-buffer run (the resident sweep stays untouched)
};

namespace {

int activate(odeintr_system_t h) {
  if (!h) return fail(ODEINTR_E_INVALID, "null system handle");
  CU_TRY(cudaSetDevice(h->device));
  return ODEINTR_OK;
}

int begin_timing(odeintr_system_t h) {
  h->last_launches = 0;
  CU_TRY(cudaEventRecord(h->ev0, h->stream));
  return ODEINTR_OK;
}
int end_timing(odeintr_system_t h) {
  CU_TRY(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  return ODEINTR_OK;
}

int launch(odeintr_system_t h, cudaKernel_t k, cudaStream_t st, unsigned grid, unsigned block, void* args_struct) {
  void* params[1] = {args_struct};
  CU_TRY(cudaLaunchKernel((const void*)k, dim3(grid), dim3(block), params, 0, st));
  h->last_launches += 1;
  g_total_launches += 1;
  return ODEINTR_OK;
}

int upload(odeintr_system_t h, void* dst, const void* src, size_t bytes) {
  if (!bytes) return ODEINTR_OK;
  CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return ODEINTR_OK;
}

// device pointer + strides for the parameter block of the current ensemble
int param_view(odeintr_system_t h, const double*& p, long long& ld, long long& inc) {
  p = nullptr; ld = 0; inc = 0;
  if (h->P == 0) return ODEINTR_OK;
  if (h->par_sets == 0) return fail(ODEINTR_E_STATE, "parameters have not been set");
  if (h->par_sets == 1) { p = h->pars.p; ld = 1; inc = 0; return ODEINTR_OK; }
  if (h->par_sets != h->m) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  p = h->pars.p; ld = (long long)h->m; inc = 1;
  return ODEINTR_OK;
}

int reserve_out(odeintr_system_t h, size_t rows) {
  const size_t need = rows * (size_t)h->N * h->m;
  cudaError_t e = h->out.reserve(need);
  if (e != cudaSuccess)
    return fail(ODEINTR_E_CUDA, std::string("cannot allocate observer record (") + std::to_string(need * 8) +
                                    " bytes): " + cudaGetErrorString(e));
  return ODEINTR_OK;
}

struct plain_schedule {
  std::vector<long long> nfull;
  std::vector<double> hrem, t;
  std::vector<unsigned char> obs;
  bool uniform = false;  // every segment nfull_uniform steps, no remainder, observe every segment
  long long n_seg = 0, nfull_uniform = 0;
  int time_rule = ODEINTR_TIME_CONST;
  long long total_steps = 0;
  double t0 = 0.0, dt = 0.0;
  bool record = false, record_final = false;
  std::vector<double> rec_t;  // the Time column the observer would have recorded
};

// integrate_const, stepper_tag: obs; do_step; time = t0 + step*dt ... final obs
void schedule_const(plain_schedule& sc, double t0, double t1, double dt, long long k, bool record) {
  const long long n = const_step_count(t0, t1, dt);
  sc.time_rule = ODEINTR_TIME_CONST;
  sc.total_steps = n;
  sc.t0 = t0; sc.dt = dt;
  sc.record = record; sc.record_final = record;
  if (!record) {
    sc.uniform = true; sc.n_seg = 1; sc.nfull_uniform = n;
    return;
  }
  sc.n_seg = (n + k - 1) / k;
  if (n % k == 0) {
    sc.uniform = true; sc.nfull_uniform = k;
  } else {
    sc.nfull.assign((size_t)sc.n_seg, k);
    sc.nfull.back() = n - k * (sc.n_seg - 1);
    sc.hrem.assign((size_t)sc.n_seg, 0.0);
    sc.obs.assign((size_t)sc.n_seg, 1);
    sc.t.resize((size_t)sc.n_seg);
    for (long long s = 0; s < sc.n_seg; ++s) sc.t[(size_t)s] = t0 + static_cast<double>(s * k) * dt;
  }
  sc.rec_t.resize((size_t)sc.n_seg + 1);
  for (long long s = 0; s < sc.n_seg; ++s) sc.rec_t[(size_t)s] = t0 + static_cast<double>(s * k) * dt;
  sc.rec_t[(size_t)sc.n_seg] = t0 + static_cast<double>(n) * dt;
}

// integrate_times, stepper_tag: per interval, steps of dt then min_abs(dt, T - t) remainders
void schedule_times(plain_schedule& sc, const double* times, size_t n_times, double dt) {
  sc.time_rule = ODEINTR_TIME_ACCUM;
  sc.t0 = times[0]; sc.dt = dt;
  sc.record = true; sc.record_final = true;
  sc.rec_t.assign(times, times + n_times);
  auto open_segment = [&](double t, bool observe) {
    sc.nfull.push_back(0); sc.hrem.push_back(0.0); sc.t.push_back(t); sc.obs.push_back(observe ? 1 : 0);
  };
  for (size_t k = 0; k + 1 < n_times; ++k) {
    double t = times[k];
    const double T = times[k + 1];
    bool first = true;
    bool open = false;  // the current segment can still take full steps
    if (!less_with_sign(t, T, dt)) {  // nothing to integrate, but the observation still happens
      open_segment(t, true);
      continue;
    }
    while (less_with_sign(t, T, dt)) {
      const double hcur = min_abs(dt, T - t);
      if (!open) { open_segment(t, first); first = false; open = true; }
      if (hcur == dt) {
        sc.nfull.back() += 1;
      } else {
        sc.hrem.back() = hcur;
        open = false;  // a remainder step closes the segment
      }
      sc.total_steps += 1;
      t += hcur;
    }
  }
  sc.n_seg = (long long)sc.nfull.size();
}

// integrate_adaptive, stepper_tag: integrate_n_steps + one step onto t1
void schedule_adaptive(plain_schedule& sc, double t0, double t1, double dt, bool record) {
  const double q = (t1 - t0) / dt;
  const long long n = q > 0 ? static_cast<long long>(q) : 0;
  const double end = t0 + static_cast<double>(n) * dt;
  const bool rem = less_with_sign(end, t1, dt);
  sc.time_rule = ODEINTR_TIME_CONST;
  sc.total_steps = n + (rem ? 1 : 0);
  sc.t0 = t0; sc.dt = dt;
  sc.record = record; sc.record_final = record;
  if (!record) {
    sc.n_seg = 1;
    if (!rem) { sc.uniform = true; sc.nfull_uniform = n; }
    else { sc.nfull.assign(1, n); sc.hrem.assign(1, t1 - end); sc.t.assign(1, t0); sc.obs.assign(1, 0); }
    return;
  }
  sc.rec_t.resize((size_t)n + 1);
  for (long long s = 0; s <= n; ++s) sc.rec_t[(size_t)s] = t0 + static_cast<double>(s) * dt;
  if (!rem) {
    sc.uniform = true; sc.n_seg = n; sc.nfull_uniform = 1;
  } else {
    sc.n_seg = n + 1;
    sc.nfull.assign((size_t)n + 1, 1); sc.nfull.back() = 0;
    sc.hrem.assign((size_t)n + 1, 0.0); sc.hrem.back() = t1 - end;
    sc.obs.assign((size_t)n + 1, 1);
    sc.t = sc.rec_t;
    sc.rec_t.push_back(t1);
  }
}

int upload_schedule(odeintr_system_t h, const plain_schedule& sc, odeintr_plain_args& a) {
  int rc;
  a.t0 = sc.t0;
  a.dt = sc.dt;
  a.n_seg = sc.n_seg;
  a.nfull_uniform = sc.nfull_uniform;
  a.time_rule = sc.time_rule;
  a.record_final = sc.record_final ? 1 : 0;
  if (sc.uniform) return ODEINTR_OK;
  const size_t n = (size_t)sc.n_seg;
  CU_TRY(h->seg_nfull.reserve(n)); CU_TRY(h->seg_hrem.reserve(n)); CU_TRY(h->seg_t.reserve(n));
  CU_TRY(h->seg_obs.reserve(n));
  // pageable sources: cudaMemcpyAsync stages them before returning, so the vectors may die afterwards
  if ((rc = upload(h, h->seg_nfull.p, sc.nfull.data(), n * sizeof(long long)))) return rc;
  if ((rc = upload(h, h->seg_hrem.p, sc.hrem.data(), n * sizeof(double)))) return rc;
  if ((rc = upload(h, h->seg_t.p, sc.t.data(), n * sizeof(double)))) return rc;
  if ((rc = upload(h, h->seg_obs.p, sc.obs.data(), n))) return rc;
  a.seg_nfull = h->seg_nfull.p; a.seg_hrem = h->seg_hrem.p; a.seg_t = h->seg_t.p; a.seg_obs = h->seg_obs.p;
  return ODEINTR_OK;
}

// Device-resident run: the whole ensemble in one launch.
int run_plain(odeintr_system_t h, const plain_schedule& sc) {
  if (!h->k_plain) return fail(ODEINTR_E_STATE, "system was not compiled with a fixed-step stepper");
  if (h->m == 0) return fail(ODEINTR_E_STATE, "Invalid initial state");
  odeintr_plain_args a;
  std::memset(&a, 0, sizeof(a));
  int rc = param_view(h, a.pars, a.par_ld, a.par_inc);
  if (rc) return rc;
  if (sc.record && (rc = reserve_out(h, sc.rec_t.size()))) return rc;
  if ((rc = upload_schedule(h, sc, a))) return rc;
  a.x = h->x.p;
  a.out = sc.record ? h->out.p : nullptr;
  a.m = (long long)h->m;
  a.ld = (long long)h->m;
  if ((rc = begin_timing(h))) return rc;
  const unsigned block = (unsigned)h->block_plain;
  const unsigned grid = (unsigned)((h->m + block - 1) / block);
  if ((rc = launch(h, h->k_plain, h->stream, grid, block, &a))) return rc;
  if ((rc = end_timing(h))) return rc;
  h->last_steps = sc.total_steps;
  if (sc.record) {
    h->rec_t = sc.rec_t;
    h->rows = sc.rec_t.size();
    h->out_m = h->m;
    h->ragged = false;
  }
  return ODEINTR_OK;
}

// Host-buffer run: the ensemble is streamed through the GPU in chunks of instances on three streams
// so the host->device copy of chunk j+1, the kernel of chunk j and the device->host copy of chunk j-1
// overlap (full-duplex PCIe + compute).  Host layouts are the C-ABI's SoA layouts with leading
// dimension n_inst; the device works on dense chunk-local copies.
int run_plain_host(odeintr_system_t h, const plain_schedule& sc, const double* x_host, size_t n_inst,
                   const double* pars_host, size_t par_sets, double* out_host, double* x_final_host, size_t chunk) {
  if (!h->k_plain) return fail(ODEINTR_E_STATE, "system was not compiled with a fixed-step stepper");
  if (!x_host || n_inst == 0) return fail(ODEINTR_E_INVALID, "Invalid initial state");
  if (h->P && (!pars_host || (par_sets != 1 && par_sets != n_inst))) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  const size_t N = (size_t)h->N, P = (size_t)h->P;
  const size_t rows = sc.record ? sc.rec_t.size() : 0;
  if (sc.record && !out_host) return fail(ODEINTR_E_INVALID, "Invalid output buffer");
  if (chunk == 0) {
    // ~1 GiB of record per chunk, at least 64 Ki instances, a multiple of the block size
    const size_t per_inst = (rows * N + 2 * N + P) * sizeof(double);
    chunk = std::max<size_t>((size_t)1 << 16, ((size_t)1 << 30) / std::max<size_t>(per_inst, 1));
  }
  chunk = std::min(chunk, n_inst);
  chunk = ((chunk + 127) / 128) * 128;
  constexpr int kSlots = 3;
  if (!h->pipe_ready) {
    for (int s = 0; s < kSlots; ++s) {
      CU_TRY(cudaStreamCreateWithFlags(&h->pipe_stream[s], cudaStreamNonBlocking));
      CU_TRY(cudaEventCreateWithFlags(&h->pipe_done[s], cudaEventDisableTiming));
    }
    h->pipe_ready = true;
  }
  for (int s = 0; s < kSlots; ++s) {
    CU_TRY(h->pipe_x[s].reserve(N * chunk));
    if (rows) CU_TRY(h->pipe_out[s].reserve(rows * N * chunk));
    if (P && par_sets == n_inst) CU_TRY(h->pipe_pars[s].reserve(P * chunk));
  }
  odeintr_plain_args base;
  std::memset(&base, 0, sizeof(base));
  int rc = upload_schedule(h, sc, base);
  if (rc) return rc;
  if (P && par_sets == 1) {
    CU_TRY(h->pipe_bcast.reserve(P));
    if ((rc = upload(h, h->pipe_bcast.p, pars_host, P * sizeof(double)))) return rc;
  }
  if ((rc = begin_timing(h))) return rc;
  // the chunk streams start after the schedule / broadcast-parameter uploads on the main stream
  cudaEvent_t start_ev = h->pipe_done[0];
  CU_TRY(cudaEventRecord(start_ev, h->stream));
  for (int s = 0; s < kSlots; ++s) CU_TRY(cudaStreamWaitEvent(h->pipe_stream[s], start_ev, 0));
  const unsigned block = (unsigned)h->block_plain;
  // Copies of earlier chunks may still be reading / writing the caller's buffers when a later enqueue fails: drain
  // every pipeline stream before reporting the failure, so the caller may free its arrays.
  auto enqueue_chunks = [&]() -> int {
    size_t j = 0;
    for (size_t i0 = 0; i0 < n_inst; i0 += chunk, ++j) {
      const int s = (int)(j % kSlots);
      cudaStream_t st = h->pipe_stream[s];
      const size_t c = std::min(chunk, n_inst - i0);
      CU_TRY(cudaMemcpy2DAsync(h->pipe_x[s].p, c * sizeof(double), x_host + i0, n_inst * sizeof(double), c * sizeof(double),
                               N, cudaMemcpyHostToDevice, st));
      odeintr_plain_args a = base;
      if (P) {
        if (par_sets == 1) { a.pars = h->pipe_bcast.p; a.par_ld = 1; a.par_inc = 0; }
        else {
          CU_TRY(cudaMemcpy2DAsync(h->pipe_pars[s].p, c * sizeof(double), pars_host + i0, n_inst * sizeof(double),
                                   c * sizeof(double), P, cudaMemcpyHostToDevice, st));
          a.pars = h->pipe_pars[s].p; a.par_ld = (long long)c; a.par_inc = 1;
        }
      }
      a.x = h->pipe_x[s].p;
      a.out = rows ? h->pipe_out[s].p : nullptr;
      a.m = (long long)c;
      a.ld = (long long)c;
      if ((rc = launch(h, h->k_plain, st, (unsigned)((c + block - 1) / block), block, &a))) return rc;
      if (rows)
        CU_TRY(cudaMemcpy2DAsync(out_host + i0, n_inst * sizeof(double), h->pipe_out[s].p, c * sizeof(double),
                                 c * sizeof(double), rows * N, cudaMemcpyDeviceToHost, st));
      if (x_final_host)
        CU_TRY(cudaMemcpy2DAsync(x_final_host + i0, n_inst * sizeof(double), h->pipe_x[s].p, c * sizeof(double),
                                 c * sizeof(double), N, cudaMemcpyDeviceToHost, st));
    }
    return ODEINTR_OK;
  };
  if ((rc = enqueue_chunks())) {
    const std::string why = g_err;
    for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(h->pipe_stream[s]);
    cudaStreamSynchronize(h->stream);
    (void)cudaGetLastError();
    return fail(rc, why);
  }
  for (int s = 0; s < kSlots; ++s) {
    CU_TRY(cudaEventRecord(h->pipe_done[s], h->pipe_stream[s]));
    CU_TRY(cudaStreamWaitEvent(h->stream, h->pipe_done[s], 0));
  }
  if ((rc = end_timing(h))) return rc;
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->last_steps = sc.total_steps;
  h->rec_t = sc.rec_t;
  h->rows = 0;  // the record lives in the caller's buffer, not on the device
  h->out_m = 0;
  return ODEINTR_OK;
}

// Large systems (compile_sys_openmp): one trajectory of N states; every Runge-Kutta stage is one fused
// launch of odeintr_lattice_stage (RHS + stage algebra), the observer is a device-to-device row copy.
struct lattice_window {      // which cells a launch computes and how local buffers map to global cells
  long long n, n_total, c0, ghost, pitch;
  int sharded;
  long long lo, hi;          // sharded run through the tiled kernels: cells the step must produce
};

int lattice_ck54_fixed_step(odeintr_system_t h, double t, double dt);  // after lattice_adaptive.h
int lattice_rkf78_fixed_step(odeintr_system_t h, double t, double dt);
int lattice_rk5_fixed_step(odeintr_system_t h, double t, double dt);
int lattice_serial_fixed_step(odeintr_system_t h, double t, double dt);

// launch descriptor of the tiled rk4 kernels for one step x_in -> x_out of the window w
void tiled_args_init(odeintr_system_t h, const lattice_window& w, const double* x_in, double* x_out, double t, double dt,
                     odeintr_lattice_tiled_args& ta) {
  std::memset(&ta, 0, sizeof(ta));
  ta.x_in = x_in; ta.x_out = x_out;
  ta.pars = h->P ? h->pars.p : nullptr;
  ta.halo = h->lattice_halo;
  ta.t = t; ta.dt = dt;
  ta.flag = h->lat_flag.p;
  ta.n = w.n;
  if (w.sharded) {
    ta.n_total = w.n_total; ta.c0 = w.c0; ta.ghost = w.ghost; ta.pitch = w.pitch; ta.sharded = 1;
  } else {
    ta.n_total = h->lattice_cells ? h->lattice_cells : w.n_total;
    ta.c0 = 0; ta.ghost = 0; ta.pitch = ta.n_total;
  }
}

int tiled_launch(odeintr_system_t h, cudaKernel_t k, unsigned grid, size_t smem, odeintr_lattice_tiled_args& ta) {
  void* params[1] = {&ta};
  CU_TRY(cudaLaunchKernel((const void*)k, dim3(grid), dim3(256), params, smem, h->stream));
  h->last_launches += 1;
  g_total_launches += 1;
  return ODEINTR_OK;
}

// general tiles (odeintr_lattice_rk4_tiled) over the cells [lo, hi) and [lo2, hi2): seam, loop ends, slab ends, ragged ends
int tiled_general(odeintr_system_t h, odeintr_lattice_tiled_args ta, long long lo, long long hi, long long lo2, long long hi2) {
  const long long tile = 256 * (long long)tile_cpt();
  const long long tiles = (hi > lo ? (hi - lo + tile - 1) / tile : 0) + (hi2 > lo2 ? (hi2 - lo2 + tile - 1) / tile : 0);
  if (tiles <= 0) return ODEINTR_OK;
  ta.lo = lo; ta.hi = hi > lo ? hi : lo; ta.lo2 = lo2; ta.hi2 = hi2;
  return tiled_launch(h, h->k_tiled, (unsigned)tiles, h->tiled_smem, ta);
}

// One rk4 step of the cells [lo, hi).  Interior cells -- the whole window of 4 * halo cells around them lies inside the
// user loop's range [glo, ghi): no periodic images, no cells the loop leaves alone -- go to the warp-autonomous
// register kernel (radius <= 2) or to the block-tiled interior kernels (wider stencils); what is left at either end
// goes to the general kernel.  Every access of every step is checked against the radius; the caller reads the flag
// when the call is over.
int tiled_produce(odeintr_system_t h, odeintr_lattice_tiled_args ta, long long lo, long long hi, long long glo,
                  long long ghi, bool allow_interior) {
  if (hi <= lo) return ODEINTR_OK;
  int rc;
  const long long tile = 256 * (long long)tile_cpt();
  const long long margin = 4 * h->lattice_halo;
  const int fields = (int)((long long)h->N / (h->lat_cells_ct > 0 ? h->lat_cells_ct : (long long)h->N));
  long long ilo = std::max(lo, glo + margin), ihi = std::min(hi, ghi - margin);
  const bool interior = allow_interior && h->lat_cells_ct == ta.n_total && ihi > ilo &&
                        !std::getenv("ODEINTR_TILED_GENERAL_ONLY");
  // 16-byte aligned windows can be moved by the TMA engine: even cell offsets, even distance between the fields
  const bool even_fields = (reinterpret_cast<uintptr_t>(ta.x_in) % 16 == 0) && (reinterpret_cast<uintptr_t>(ta.x_out) % 16 == 0) &&
                           (fields == 1 || ta.pitch % 2 == 0);
  int mode = 0;  // 1 = warp kernel, 2 = block-tiled interior kernels
  if (interior && h->k_warp && h->warp_blocks > 0 && !std::getenv("ODEINTR_NO_WARP")) {
    const long long P = 32LL * h->warp_cpt - 8LL * h->warp_halo;
    // the warp kernel evaluates EVERY cell of its windows, the outermost (discarded) ones included: keep their own
    // neighbours inside the user loop's range as well, so that an absolute index never reaches across the seam
    long long wlo = std::max(ilo, glo + margin + h->lattice_halo), whi = std::min(ihi, ghi - margin - h->lattice_halo);
    const bool tma = even_fields && !std::getenv("ODEINTR_NO_TMA");
    if (tma) {
      if ((wlo - ta.c0) & 1) ++wlo;
      if ((whi - ta.c0) & 1) --whi;
    }
    if (whi - wlo >= P) {
      mode = 1; ilo = wlo; ihi = whi; ta.tma = tma ? 1 : 0;
    }
  }
  long long n_full = 0;
  if (!mode && interior && h->k_tiled_int && 8 * h->lattice_halo <= 256) {
    n_full = (ihi - ilo) / tile;
    if (n_full > 0) { mode = 2; ihi = ilo + n_full * tile; }
  }
  if (!mode) ilo = ihi = hi;
  if ((rc = tiled_general(h, ta, lo, ilo, ihi, hi))) return rc;
  if (mode == 1) {
    const long long P = 32LL * h->warp_cpt - 8LL * h->warp_halo;
    const long long warp_tiles = (ihi - ilo + P - 1) / P;
    ta.lo = ilo; ta.hi = ihi; ta.sharded = 0;  // the kernel addresses both layouts through (pitch, c0)
    const long long slots = (long long)h->sm_count * h->warp_blocks;
    const unsigned grid = (unsigned)std::min<long long>((warp_tiles + 7) / 8, slots);
    return tiled_launch(h, h->k_warp, grid, h->warp_smem, ta);
  }
  if (mode == 2) {
    ta.lo = ilo; ta.hi = ihi; ta.sharded = 0;
    ta.tile_first = 0; ta.tile_skip = (int)n_full;
    // the persistent kernel prefetches the next tile's window with the TMA engine while it computes the current one
    const bool aligned = even_fields && ((ilo - ta.c0) % 2 == 0);
    if (h->tiled_tma_blocks > 0 && aligned) {
      const long long slots = (long long)h->sm_count * h->tiled_tma_blocks;
      return tiled_launch(h, h->k_tiled_tma, (unsigned)std::min<long long>(n_full, slots), h->tiled_tma_smem, ta);
    }
    return tiled_launch(h, h->k_tiled_int, (unsigned)n_full, h->tiled_int_smem, ta);
  }
  return ODEINTR_OK;
}

// Two rk4 steps of a 1-D large system in ONE pass of the warp kernel (a.nsub = 2): the interior goes x -> x'' without
// the intermediate state ever reaching HBM (8 B per cell and step instead of 16); the few hundred cells at the ends of
// the loop range / the periodic seam take the general kernel twice, through an intermediate array.  Pays when a single
// step is bound by HBM (several fields).  Same operations per cell as two separate steps: bit-identical.
enum { LATTICE_NOT_FUSED = -1001 };
int lattice_step_pair(odeintr_system_t h, const lattice_window& w, double t, double t2, double dt) {
  const char* knob = std::getenv("ODEINTR_FUSE");
  const int fields = h->lat_cells_ct > 0 ? (int)((long long)h->N / h->lat_cells_ct) : 1;
  // measured (profiles/r02_lattice_warp.md): two fields 0.74 -> 0.61 ms per step (HBM-bound otherwise), one field
  // 0.90 -> 0.85: on by default; ODEINTR_FUSE=0 switches it off for A/B measurements
  const bool want = knob ? std::atoi(knob) != 0 : true;
  (void)fields;
  if (!want || w.sharded || h->lattice_stages != 4 || !h->tiled_on || h->tiled2d_on || !h->k_tiled || !h->k_warp ||
      h->warp_blocks <= 0 || h->k_serial || std::getenv("ODEINTR_NO_WARP") || std::getenv("ODEINTR_TILED_GENERAL_ONLY"))
    return LATTICE_NOT_FUSED;
  odeintr_lattice_tiled_args ta;
  tiled_args_init(h, w, h->lat_x, h->lat_a, t, dt, ta);
  if (h->lat_cells_ct != ta.n_total) return LATTICE_NOT_FUSED;
  const long long glo = h->lat_start > 0 ? h->lat_start : 0;
  const long long ghi = h->lat_bound < ta.n_total ? h->lat_bound : ta.n_total;
  const long long m1 = 4 * h->lattice_halo;  // margin of one step
  const long long P = 32LL * h->warp_cpt - 16LL * h->warp_halo;
  long long ilo = glo + 2 * m1 + h->lattice_halo, ihi = ghi - 2 * m1 - h->lattice_halo;
  const bool tma = (reinterpret_cast<uintptr_t>(ta.x_in) % 16 == 0) && (reinterpret_cast<uintptr_t>(ta.x_out) % 16 == 0) &&
                   (fields == 1 || ta.pitch % 2 == 0) && !std::getenv("ODEINTR_NO_TMA");
  if (tma) {
    if (ilo & 1) ++ilo;
    if (ihi & 1) --ihi;
  }
  if (P < 32 || ihi - ilo < P) return LATTICE_NOT_FUSED;
  const size_t N = (size_t)h->N;
  CU_TRY(h->lat_buf_acc.reserve(N));
  double* mid = h->lat_buf_acc.p;
  const bool covers_all = h->lat_cells_ct > 0 && h->lat_start <= 0 && h->lat_bound >= h->lat_cells_ct;
  if (!covers_all && !h->lat_mid_primed)  // cells no loop writes never change: one copy per integrate call
    CU_TRY(cudaMemcpyAsync(mid, h->lat_x, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->lat_mid_primed = true;
  int rc;
  // the ends, first step: x -> mid on the end ranges widened by one step's margin (what the second step reads)
  odeintr_lattice_tiled_args e1 = ta;
  e1.x_out = mid;
  if ((rc = tiled_general(h, e1, glo, std::min(ilo + m1, ghi), std::max(ihi - m1, glo), ghi))) return rc;
  // the ends, second step: mid -> x_out
  odeintr_lattice_tiled_args e2 = ta;
  e2.x_in = mid; e2.t = t2;
  if ((rc = tiled_general(h, e2, glo, ilo, ihi, ghi))) return rc;
  // the interior: both steps in registers
  ta.lo = ilo; ta.hi = ihi; ta.sharded = 0; ta.tma = tma ? 1 : 0; ta.nsub = 2; ta.t2 = t2;
  const long long warp_tiles = (ihi - ilo + P - 1) / P;
  const unsigned grid = (unsigned)std::min<long long>((warp_tiles + 7) / 8, (long long)h->sm_count * h->warp_blocks);
  if ((rc = tiled_launch(h, h->k_warp, grid, h->warp_smem, ta))) return rc;
  std::swap(h->lat_x, h->lat_a);
  return ODEINTR_OK;
}

int lattice_step_halo(odeintr_system_t h, const lattice_window& w, double t, double dt, int halo) {
  // sharded runs compute a window that shrinks by the stencil radius `halo` with every stage, so one halo
  // exchange of width halo * stages per step is enough (redundant work on 2 * halo * stages * (stages-1)/2 cells)
  odeintr_lattice_args a;
  std::memset(&a, 0, sizeof(a));
  a.pars = h->P ? h->pars.p : nullptr;
  a.n = w.n; a.n_total = w.n_total; a.c0 = w.c0; a.ghost = w.ghost; a.pitch = w.pitch;
  double* x = h->lat_x;
  double* A = h->lat_a;
  double* B = h->lat_b;
  double* acc = h->lat_acc;
  const unsigned grid = (unsigned)h->sm_count * 8u, block = 256u;
  int rc;
  if (h->k_serial && (h->lattice_stages == 4 || h->lattice_stages == 1)) {
    if (w.sharded) return fail(ODEINTR_E_STATE, "general (non-cell-loop) snippets cannot be sharded");
    return lattice_serial_fixed_step(h, t, dt);
  }
  if (h->lattice_stages == 4 && h->tiled2d_on && h->k_tiled2d && !w.sharded) {
    // M x M lattice with a short reach: the whole rk4 step in one temporally blocked pass over 32 x 64 tiles
    odeintr_lattice_tiled_args ta;
    std::memset(&ta, 0, sizeof(ta));
    ta.x_in = x; ta.x_out = A;
    ta.pars = a.pars;
    ta.n = w.n; ta.n_total = w.n_total;
    ta.halo = h->lattice_halo;
    ta.t = t; ta.dt = dt;
    ta.flag = h->lat_flag.p;
    const long long side = h->lat_side;
    if (h->stream2d_blocks > 0 && h->lattice_halo == 1) {
      // strips of 56 columns x segments of stream2d_rows rows, one warp each, strips fastest (neighbouring warps walk
      // down neighbouring strips: whole rows of the lattice are touched together)
      const long long tasks = ((side + 55) / 56) * ((side + h->stream2d_rows - 1) / h->stream2d_rows);
      unsigned sgrid = (unsigned)std::min<long long>((tasks + 7) / 8, (long long)h->sm_count * h->stream2d_blocks);
      // (tests: a grid of a block or two makes every warp walk through several tasks, as on lattices with more tasks
      //  than resident warps)
      if (const char* c = std::getenv("ODEINTR_STREAM_MAX_BLOCKS")) sgrid = std::max(1u, std::min(sgrid, (unsigned)std::atoi(c)));
      void* sparams[1] = {&ta};
      CU_TRY(cudaLaunchKernel((const void*)h->k_stream2d, dim3(sgrid), dim3(256), sparams, 0, h->stream));
      h->last_launches += 1;
      g_total_launches += 1;
      std::swap(h->lat_x, h->lat_a);
      return ODEINTR_OK;
    }
    const long long th = tile2_rows();
    const unsigned tiles = (unsigned)(((side + 63) / 64) * ((side + th - 1) / th));
    void* params[1] = {&ta};
    CU_TRY(cudaLaunchKernel((const void*)h->k_tiled2d, dim3(tiles), dim3(256), params, h->tiled2d_smem, h->stream));
    h->last_launches += 1;
    g_total_launches += 1;
    std::swap(h->lat_x, h->lat_a);
    return ODEINTR_OK;
  }
  if (h->lattice_stages == 4 && h->k_tiled && (w.sharded ? h->sh_tiled : h->tiled_on)) {
    // known stencil radius: the whole rk4 step in one temporally blocked pass, x -> A, then the caller lets the
    // buffers trade places (16 B per cell per step instead of 128)
    odeintr_lattice_tiled_args ta;
    tiled_args_init(h, w, x, A, t, dt, ta);
    // cells of the user's loop [glo, ghi), and the cells this launch produces [lo, hi)
    const long long glo = h->lat_start > 0 ? h->lat_start : 0;
    const long long ghi = h->lat_bound < ta.n_total ? h->lat_bound : ta.n_total;
    if ((rc = tiled_produce(h, ta, w.sharded ? w.lo : glo, w.sharded ? w.hi : ghi, glo, ghi, true))) return rc;
    if (!w.sharded) std::swap(h->lat_x, h->lat_a);
    return ODEINTR_OK;
  }
  auto stage = [&](int s, int n_stages, const double* xs, double* xt, const double* acc_in, double* acc_out,
                   double ts, double a_dt, double b_dt) {
    a.xs = xs; a.x = x; a.xt = xt; a.acc_in = acc_in; a.acc_out = acc_out; a.t = ts; a.a_dt = a_dt; a.b_dt = b_dt;
    const long long margin = w.sharded ? (long long)halo * (n_stages - s) : 0;
    a.lo = w.c0 - margin;
    a.hi = w.c0 + w.n + margin;
    a.halo = halo;
    return launch(h, w.sharded ? h->k_lattice_sharded : h->k_lattice, h->stream, grid, block, &a);
  };
  if (h->lattice_stages == 6) return lattice_ck54_fixed_step(h, t, dt);
  if (h->lattice_stages == 13) return lattice_rkf78_fixed_step(h, t, dt);
  if (h->lattice_stages == 7) return lattice_rk5_fixed_step(h, t, dt);
  if (h->lattice_stages == 4) {
    // runge_kutta4 = explicit_generic_rk<4,4>: a = {1/2}, {0,1/2}, {0,0,1}; b = {1/6,1/3,1/3,1/6}
    const double half = 1.0 / 2.0, b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
    if ((rc = stage(1, 4, x, A, x, acc, t, half * dt, b1 * dt))) return rc;
    if ((rc = stage(2, 4, A, B, acc, acc, t + half * dt, half * dt, b2 * dt))) return rc;
    if ((rc = stage(3, 4, B, A, acc, acc, t + half * dt, 1.0 * dt, b2 * dt))) return rc;
    if ((rc = stage(4, 4, A, nullptr, acc, x, t + 1.0 * dt, 0.0, b1 * dt))) return rc;
  } else {
    // euler: x_new = x + dt * f(x); written to A, then the buffers trade places
    if ((rc = stage(1, 1, x, nullptr, x, A, t, 0.0, dt))) return rc;
    std::swap(h->lat_x, h->lat_a);
  }
  return ODEINTR_OK;
}

int lattice_step(odeintr_system_t h, const lattice_window& w, double t, double dt, int) {
  return lattice_step_halo(h, w, t, dt, 0);
}

int lattice_prepare(odeintr_system_t h, size_t total) {
  // Work buffers of a fixed-step run, only what the selected path uses.  Stage-per-launch rk4: two stage states and
  // the running sum, all starting as copies of x so that cells no loop writes keep their value in every stage.
  // Tiled rk4 (and euler): ONE array the state trades places with; it needs the copy only when the user's loop leaves
  // cells alone.  A radius measured by the probe may turn out too short during the run (indices that depend on the
  // state): the second array then keeps the initial state for the repeat on the staged kernels.
  const bool tiled = h->lattice_stages == 4 && ((h->tiled_on && h->k_tiled) || (h->tiled2d_on && h->k_tiled2d));
  const bool euler = h->lattice_stages == 1;
  const bool covers_all = h->lat_cells_ct > 0 && h->lat_start <= 0 && h->lat_bound >= h->lat_cells_ct;
  const bool need_b = tiled ? h->tiled_auto : !euler;
  const bool need_acc = !tiled && !euler;
  CU_TRY(h->lat_buf_a.reserve(total));
  if (!(tiled && covers_all))
    CU_TRY(cudaMemcpyAsync(h->lat_buf_a.p, h->x.p, total * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (need_b) {
    CU_TRY(h->lat_buf_b.reserve(total));
    CU_TRY(cudaMemcpyAsync(h->lat_buf_b.p, h->x.p, total * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  if (need_acc) {
    CU_TRY(h->lat_buf_acc.reserve(total));
    CU_TRY(cudaMemcpyAsync(h->lat_buf_acc.p, h->x.p, total * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  return ODEINTR_OK;
}

// shared-memory budgets of the two tiled kernels for a stencil radius; false when no tile can hold it
bool tiled_configure(odeintr_system_t h, long long halo, int fields) {
  if (!h->k_tiled || halo < 0 || halo > 4096) return false;
  // the window of a tile wraps around the lattice at most once: tiny lattices stay on the staged kernels
  if ((long long)h->N / fields < 8 * halo + 2) return false;
  const size_t width = 256 * (size_t)tile_cpt() + 8 * (size_t)halo;
  const size_t bytes = 4 * (size_t)fields * width * sizeof(double);
  if (bytes > 200 * 1024) return false;
  if (cudaFuncSetAttribute((const void*)h->k_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess ||
      (h->k_tiled_int && cudaFuncSetAttribute((const void*)h->k_tiled_int, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(bytes / 2)) != cudaSuccess)) {
    (void)cudaGetLastError();
    return false;
  }
  h->tiled_smem = bytes;
  h->tiled_int_smem = bytes / 2;
  h->lattice_halo = halo;
  h->tiled_tma_blocks = 0;
  h->tiled_tma_smem = bytes + 16;
  if (h->k_tiled_tma && h->k_tiled_int && !std::getenv("ODEINTR_NO_TMA") &&
      cudaFuncSetAttribute((const void*)h->k_tiled_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)h->tiled_tma_smem) == cudaSuccess) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)h->k_tiled_tma, 256, h->tiled_tma_smem) == cudaSuccess)
      h->tiled_tma_blocks = nb;
  }
  (void)cudaGetLastError();
  // radius <= 2: the warp-autonomous register kernel takes the interior cells (lattice_warp.cuh)
  h->k_warp = nullptr; h->warp_blocks = 0;
  const int wh = halo < 1 ? 1 : (int)halo;
  if (wh <= 2 && h->k_warp_h[wh] && h->warp_cpt >= wh && (h->warp_cpt & 1)) {
    const size_t wsmem = 8 * (4 * (size_t)fields * 32 * (size_t)h->warp_cpt * sizeof(double) + 2 * sizeof(unsigned long long));
    int nb = 0;
    if (wsmem <= 220 * 1024 &&
        cudaFuncSetAttribute((const void*)h->k_warp_h[wh], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)h->k_warp_h[wh], 256, wsmem) == cudaSuccess && nb > 0) {
      h->k_warp = h->k_warp_h[wh];
      h->warp_halo = wh;
      h->warp_smem = wsmem;
      h->warp_blocks = nb;
    }
    (void)cudaGetLastError();
  }
  return true;
}

// M x M lattices: shared-memory budget of odeintr_lattice_rk4_tiled2d (two windows of (32 + 8 halo) x (64 + 8 halo))
bool tiled2d_configure(odeintr_system_t h, long long halo) {
  if (h->lat_side <= 0 || halo < 0 || halo > 2 || h->lat_side < 8) return false;
  if (halo == 0) halo = 1;  // no coupling at all: the radius-1 kernel does
  if (!h->k_tiled2d_h[halo]) return false;
  h->k_tiled2d = h->k_tiled2d_h[halo];
  const size_t bytes = 2 * (size_t)(tile2_rows() + 8 * halo) * (size_t)(64 + 8 * halo) * sizeof(double);
  if (cudaFuncSetAttribute((const void*)h->k_tiled2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  h->tiled2d_smem = bytes;
  h->lattice_halo = halo;
  // radius 1, even M, a loop over all cells: the streaming kernel (lattice_stream2d.cuh) takes the step
  h->stream2d_blocks = 0;
  const bool covers_all = h->lat_start <= 0 && h->lat_bound >= (long long)h->N;
  if (halo == 1 && h->k_stream2d && h->k_stream2d_info && h->lat_side % 2 == 0 && h->lat_side >= 64 && covers_all &&
      !std::getenv("ODEINTR_NO_STREAM2D")) {
    if (h->lat_flag.reserve(10) == cudaSuccess) {
      void* wout = h->lat_flag.p;
      void* wparams[1] = {&wout};
      int rows = 0, nb = 0;
      if (cudaLaunchKernel((const void*)h->k_stream2d_info, dim3(1), dim3(1), wparams, 0, h->stream) == cudaSuccess &&
          cudaMemcpyAsync(&rows, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream) == cudaSuccess &&
          cudaStreamSynchronize(h->stream) == cudaSuccess && rows > 0 &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)h->k_stream2d, 256, 0) == cudaSuccess && nb > 0) {
        h->stream2d_rows = rows;
        h->stream2d_blocks = nb;
      }
      (void)cudaGetLastError();
    }
  }
  return true;
}

// a slab of a sharded run can step through the tiled kernels: rk4, radius a tile can hold
bool slab_tiled_possible(odeintr_system_t h, int halo) {
  return h->k_tiled && h->lattice_stages == 4 && halo >= 1 && halo <= 32 && h->lat_cells_ct > 0 &&
         h->lat_cells_ct < (1LL << 31) && !std::getenv("ODEINTR_NO_TILED");
}

// odeintr_lattice_info: loop range, cells per field, layout of the dxdt targets
int lattice_query_info(odeintr_system_t h) {
  h->lat_start = 0;
  h->lat_bound = h->lat_cells_ct = 0;
  if (!h->k_info) return ODEINTR_OK;
  CU_TRY(h->lat_flag.reserve(10));  // 40 bytes: five long long
  void* out = h->lat_flag.p;
  void* params[1] = {&out};
  CU_TRY(cudaLaunchKernel((const void*)h->k_info, dim3(1), dim3(1), params, 0, h->stream));
  long long info[5] = {0, 0, 0, 0, 0};
  CU_TRY(cudaMemcpyAsync(info, h->lat_flag.p, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (h->k_warp_info) {  // cells per lane of the warp kernels
    void* wout = h->lat_flag.p;
    void* wparams[1] = {&wout};
    int cpt = 0;
    CU_TRY(cudaLaunchKernel((const void*)h->k_warp_info, dim3(1), dim3(1), wparams, 0, h->stream));
    CU_TRY(cudaMemcpyAsync(&cpt, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->warp_cpt = cpt;
  }
  h->lat_start = info[0];
  h->lat_bound = info[1];
  h->lat_cells_ct = info[2];
  h->lat_side = info[4];
  if (!info[3])
    return fail(ODEINTR_E_INVALID, "large systems must write dxdt[loop variable + constant] "
                                   "(one contiguous field per written entry)");
  return ODEINTR_OK;
}

// Measure the reach of the loop body (odeintr_lattice_probe) and switch the tiled kernels on when a tile can hold it.
int tiled_autodetect(odeintr_system_t h) {
  // the probe evaluates the snippet with the system's parameters: wait until they are in place (the integrate call
  // that follows reports "Invalid parameter vector" if they never arrive)
  if (h->P && !h->pars.p) return ODEINTR_OK;
  h->tiled_pending = false;
  h->tiled_on = h->tiled2d_on = false;
  if (!h->k_probe || !(h->k_tiled || h->k_tiled2d) || h->lattice_stages != 4 || h->lat_cells_ct <= 0 ||
      h->lat_cells_ct >= (1LL << 31) || std::getenv("ODEINTR_NO_TILED"))
    return ODEINTR_OK;
  const int fields = (int)((long long)h->N / h->lat_cells_ct);
  CU_TRY(h->lat_flag.reserve(8));
  CU_TRY(cudaMemsetAsync(h->lat_flag.p, 0, sizeof(int), h->stream));
  const double* pars = h->P ? h->pars.p : nullptr;
  double t = 0.0;
  int* out = h->lat_flag.p;
  void* params[3] = {&pars, &t, &out};
  CU_TRY(cudaLaunchKernel((const void*)h->k_probe, dim3((unsigned)h->sm_count * 8u), dim3(256), params, 0, h->stream));
  int reach = 0;
  CU_TRY(cudaMemcpyAsync(&reach, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  // beyond 32 cells the margins (4 * halo per side, recomputed by every tile) stop paying for themselves, and the
  // interior kernel's one-margin-cell-per-thread layout ends
  if (h->lat_side > 0) {  // M x M lattice: the probe reports the larger of the row and column reach
    if (!tiled2d_configure(h, reach)) return ODEINTR_OK;
    h->tiled2d_on = true;
    h->tiled_auto = true;
    return ODEINTR_OK;
  }
  if (reach > 32 || !tiled_configure(h, reach, fields)) return ODEINTR_OK;
  h->lattice_cells = h->lat_cells_ct;
  h->tiled_on = true;
  h->tiled_auto = true;
  return ODEINTR_OK;
}

enum { LATTICE_REACH_EXCEEDED = -1000 };  // internal: a tiled step read further than the staged radius

int run_lattice_once(odeintr_system_t h, const plain_schedule& sc) {
  if (h->m != 1) return fail(ODEINTR_E_STATE, "Invalid initial state");
  const size_t N = (size_t)h->N;
  int rc;
  if (h->P && h->par_sets != 1) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  if ((rc = lattice_prepare(h, N))) return rc;
  h->lat_x = h->x.p; h->lat_a = h->lat_buf_a.p; h->lat_b = h->lat_buf_b.p; h->lat_acc = h->lat_buf_acc.p;
  if (sc.record && (rc = reserve_out(h, sc.rec_t.size()))) return rc;
  lattice_window w = {(long long)N, (long long)N, 0, 0, 0, 0, 0, 0};
  CU_TRY(h->lat_flag.reserve(1));
  CU_TRY(cudaMemsetAsync(h->lat_flag.p, 0, sizeof(int), h->stream));
  h->lat_work_primed = false;
  h->lat_mid_primed = false;
  if ((rc = begin_timing(h))) return rc;
  size_t row = 0;
  auto observe = [&]() -> int {
    CU_TRY(cudaMemcpyAsync(h->out.p + row * N, h->lat_x, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    ++row;
    return ODEINTR_OK;
  };
  double t = sc.t0;
  long long step = 0;
  for (long long s = 0; s < sc.n_seg; ++s) {
    if (!sc.uniform) t = sc.t[(size_t)s];
    if (sc.record && (sc.uniform || sc.obs[(size_t)s])) if ((rc = observe())) return rc;
    const long long nfull = sc.uniform ? sc.nfull_uniform : sc.nfull[(size_t)s];
    for (long long k = 0; k < nfull;) {
      if (k + 1 < nfull) {  // two steps in one pass where that pays (lattice_step_pair)
        const double t_b = (sc.time_rule == ODEINTR_TIME_CONST) ? sc.t0 + static_cast<double>(step + 1) * sc.dt : t + sc.dt;
        rc = lattice_step_pair(h, w, t, t_b, sc.dt);
        if (rc == ODEINTR_OK) {
          step += 2; k += 2;
          t = (sc.time_rule == ODEINTR_TIME_CONST) ? sc.t0 + static_cast<double>(step) * sc.dt : t_b + sc.dt;
          continue;
        }
        if (rc != LATTICE_NOT_FUSED) return rc;
      }
      if ((rc = lattice_step(h, w, t, sc.dt, 0))) return rc;
      ++step; ++k;
      t = (sc.time_rule == ODEINTR_TIME_CONST) ? sc.t0 + static_cast<double>(step) * sc.dt : t + sc.dt;
    }
    if (!sc.uniform && sc.hrem[(size_t)s] != 0.0) {
      const double hr = sc.hrem[(size_t)s];
      if ((rc = lattice_step(h, w, t, hr, 0))) return rc;
      t += hr;
    }
  }
  if (sc.record && sc.record_final) if ((rc = observe())) return rc;
  if (h->lat_x != h->x.p)  // euler swapped an odd number of times: bring the state home
    CU_TRY(cudaMemcpyAsync(h->x.p, h->lat_x, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if ((rc = end_timing(h))) return rc;
  h->last_steps = sc.total_steps;
  if (sc.record) {
    h->rec_t = sc.rec_t;
    h->rows = sc.rec_t.size();
    h->out_m = 1;
    h->ragged = false;
  }
  if (h->tiled_on || h->tiled2d_on) {
    int flag = 0;
    CU_TRY(cudaMemcpyAsync(&flag, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    if (flag) return LATTICE_REACH_EXCEEDED;
  }
  return ODEINTR_OK;
}

// Ensemble of wider systems (m > 1 trajectories of a compile_sys_openmp system): one block per trajectory, the
// whole integrate call in one launch of odeintr_lattice_ensemble (lattice_ensemble.cuh).
int run_lattice_ensemble(odeintr_system_t h, const plain_schedule& sc) {
  // rk4: the register-blocked kernel of lattice_ensemble.cuh; euler / rk54 / rk5 / rk78: the plain steppers on a
  // distributed state (lattice_ensemble_adaptive.cuh)
  const bool other = h->lattice_stages != 4;
  cudaKernel_t kern = other ? h->k_lat_ens_plain : h->k_lat_ens;
  if (!kern)
    return fail(ODEINTR_E_STATE, "ensembles of large systems are provided for 1-D lattices of at most 8192 / fields cells per "
                                 "field (rk4) or 2048 cells and 4 states per thread (the other methods); run the trajectories "
                                 "one at a time");
  if (h->lat_cells_ct <= 0) return fail(ODEINTR_E_STATE, "system was not compiled with compile_sys_openmp");
  if (h->P && !h->pars.p) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  int rc;
  // stencil radius: declared, or measured by the probe (parameters of instance 0; index expressions that depend on
  // parameters or state are caught by the per-access check)
  long long halo = h->lattice_halo;
  if (other) {
    if (!h->k_probe) return fail(ODEINTR_E_STATE, "the generated translation unit has no probe kernel");
    int reach = 0;
    CU_TRY(h->lat_flag.reserve(10));
    CU_TRY(cudaMemsetAsync(h->lat_flag.p, 0, sizeof(int), h->stream));
    const double* ppars = h->P ? h->pars.p : nullptr;
    double pt = 0.0;
    int* pout = h->lat_flag.p;
    void* pparams[3] = {&ppars, &pt, &pout};
    CU_TRY(cudaLaunchKernel((const void*)h->k_probe, dim3((unsigned)h->sm_count * 8u), dim3(256), pparams, 0, h->stream));
    CU_TRY(cudaMemcpyAsync(&reach, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    if (reach > 32) return fail(ODEINTR_E_INVALID, "ensembles of large systems need a stencil radius of at most 32 cells");
    halo = reach < 1 ? 1 : reach;
  } else if (!h->tiled_on) {
    const bool pending = h->tiled_pending, was_auto = h->tiled_auto;
    if ((rc = tiled_autodetect(h))) return rc;  // sets lattice_halo when a tile can hold the reach
    if (!h->tiled_on) {
      h->tiled_pending = pending; h->tiled_auto = was_auto;
      return fail(ODEINTR_E_INVALID, "ensembles of large systems need a stencil radius of at most 32 cells");
    }
    halo = h->lattice_halo;
  }
  const long long L = h->lat_cells_ct;
  const int fields = (int)((long long)h->N / L);
  if (2 * halo + 1 > L) return fail(ODEINTR_E_INVALID, "stencil radius too wide for the lattice");
  odeintr_lattice_ensemble_args a;
  std::memset(&a, 0, sizeof(a));
  if ((rc = param_view(h, a.p.pars, a.p.par_ld, a.p.par_inc))) return rc;
  if (sc.record && (rc = reserve_out(h, sc.rec_t.size()))) return rc;
  if ((rc = upload_schedule(h, sc, a.p))) return rc;
  a.p.x = h->x.p;
  a.p.out = sc.record ? h->out.p : nullptr;
  a.p.m = (long long)h->m;
  a.p.ld = (long long)h->m;
  a.first = 0;
  a.halo = halo;
  CU_TRY(h->lat_flag.reserve(8));
  CU_TRY(cudaMemsetAsync(h->lat_flag.p, 0, sizeof(int), h->stream));
  a.flag = h->lat_flag.p;
  // threads per trajectory and trajectories per block, as lattice_ensemble.cuh derives them from the lattice length
  const unsigned nt = (unsigned)(L > 2048 ? 512 : (L >= 256 ? 256 : (L + 31) / 32 * 32));
  const unsigned group = nt == 32 ? 4 : 1;
  const size_t smem = group * (2 * (size_t)fields * (size_t)(L + 2 * halo) + (other ? 2 * (nt / 32) : 0)) * sizeof(double);
  if (smem > 200 * 1024) return fail(ODEINTR_E_INVALID, "lattice too large for one block's shared memory");
  CU_TRY(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned block = nt * group;
  const unsigned grid = (unsigned)((h->m + group - 1) / group);
  if ((rc = begin_timing(h))) return rc;
  void* params[1] = {&a};
  CU_TRY(cudaLaunchKernel((const void*)kern, dim3(grid), dim3(block), params, smem, h->stream));
  h->last_launches += 1;
  g_total_launches += 1;
  if ((rc = end_timing(h))) return rc;
  h->last_steps = sc.total_steps;
  if (sc.record) {
    h->rec_t = sc.rec_t;
    h->rows = sc.rec_t.size();
    h->out_m = h->m;
    h->ragged = false;
  }
  int flag = 0;
  CU_TRY(cudaMemcpyAsync(&flag, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (flag)
    return fail(ODEINTR_E_INVALID, "the system definition reads x further than the stencil radius (" +
                                       std::to_string(halo) + " cells) from the cell it computes");
  return ODEINTR_OK;
}

int run_lattice(odeintr_system_t h, const plain_schedule& sc) {
  int rc;
  if (h->m > 1) return run_lattice_ensemble(h, sc);
  if (h->tiled_pending && (rc = tiled_autodetect(h))) return rc;
  rc = run_lattice_once(h, sc);
  if (rc != LATTICE_REACH_EXCEEDED) return rc;
  if (!h->tiled_auto)
    return fail(ODEINTR_E_INVALID, "the system definition reads x further than the declared halo (" +
                                       std::to_string(h->lattice_halo) + " cells) from the cell it computes");
  // The probe saw a shorter reach than the run needed (the snippet's indices depend on the state or on time):
  // the staged kernels have no such limit.  lattice_prepare left the initial state in the second stage buffer,
  // which the tiled kernels never touch.
  h->tiled_on = h->tiled2d_on = false;
  CU_TRY(cudaMemcpyAsync(h->x.p, h->lat_buf_b.p, (size_t)h->N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return run_lattice_once(h, sc);
}

// cells per thread of the tiled dopri5 kernel (ODEINTR_TILE5_CPT in lattice_dopri5_tiled.cuh)
int tile5_cpt(int fields) {
  const char* c = std::getenv("ODEINTR_TILE5_CPT");
  const int v = c ? std::atoi(c) : (fields > 1 ? 2 : 4);
  return v > 0 ? v : 1;
}

// Decide once per system whether dopri5 attempts run through the temporally blocked kernel: measure the reach of
// the loop body (odeintr_lattice_probe) and check that a tile with 6 * halo margin cells per side fits.
int tiled5_decide(odeintr_system_t h) {
  if (h->tiled5 >= 0) return ODEINTR_OK;
  if (h->P && !h->pars.p) return ODEINTR_OK;  // no parameters yet: the probe would have nothing to read
  h->tiled5 = 0;
  if (h->lat_side > 0) {
    // M x M lattice (one field): radius 1, even M >= 64, a loop over all cells -> the streaming attempt kernel
    const bool covers_all = h->lat_start <= 0 && h->lat_bound >= (long long)h->N;
    if (!h->k_stream5 || !h->k_probe || h->lat_side % 2 != 0 || h->lat_side < 64 || !covers_all ||
        h->lat_cells_ct != (long long)h->N || h->lat_cells_ct >= (1LL << 31) || std::getenv("ODEINTR_NO_TILED") ||
        std::getenv("ODEINTR_NO_STREAM2D"))
      return ODEINTR_OK;
    CU_TRY(h->lat_flag.reserve(10));
    CU_TRY(cudaMemsetAsync(h->lat_flag.p, 0, sizeof(int), h->stream));
    const double* pars = h->P ? h->pars.p : nullptr;
    double t = 0.0;
    int* out = h->lat_flag.p;
    void* params[3] = {&pars, &t, &out};
    CU_TRY(cudaLaunchKernel((const void*)h->k_probe, dim3((unsigned)h->sm_count * 8u), dim3(256), params, 0, h->stream));
    int reach = 0, nb = 0;
    CU_TRY(cudaMemcpyAsync(&reach, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    if (reach > 1) return ODEINTR_OK;  // (the probe reports the larger of the row and column reach)
    int threads = 256;
    if (h->k_stream5_info) {  // threads per block the kernel was compiled for
      void* wout = h->lat_flag.p;
      void* wparams[1] = {&wout};
      if (cudaLaunchKernel((const void*)h->k_stream5_info, dim3(1), dim3(1), wparams, 0, h->stream) == cudaSuccess &&
          cudaMemcpyAsync(&threads, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream) == cudaSuccess)
        cudaStreamSynchronize(h->stream);
      (void)cudaGetLastError();
      if (threads < 32 || threads > 1024 || threads % 32) threads = 256;
    }
    h->stream5_threads = threads;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)h->k_stream5, threads, 0) != cudaSuccess || nb <= 0) {
      (void)cudaGetLastError();
      return ODEINTR_OK;
    }
    // rows per warp task: the fewest waves of tasks (strips x segments) over the resident warps with segments of at
    // most 512 rows (12 rows of pipeline fill each), so that the last wave is a full one
    const long long strips = (h->lat_side + 51) / 52, slots = (long long)h->sm_count * nb * (threads / 32);
    long long rows = 256;
    if (const char* c = std::getenv("ODEINTR_STREAM5_ROWS")) rows = std::atoll(c);
    else
      for (long long wv = 1; wv <= 64; ++wv) {
        const long long segs = wv * slots / strips;
        if (segs < 1) continue;
        const long long r = (h->lat_side + segs - 1) / segs;
        if (r <= 512) { rows = r < 48 ? 48 : r; break; }
      }
    h->stream5_blocks = nb; h->stream5_rows = (int)rows;
    h->tiled5_halo = 1;
    h->tiled5 = 2;
    return ODEINTR_OK;
  }
  if (!h->k_dopri5_tiled || !h->k_probe || h->lat_cells_ct <= 0 || h->lat_cells_ct >= (1LL << 31) ||
      std::getenv("ODEINTR_NO_TILED"))
    return ODEINTR_OK;
  CU_TRY(h->lat_flag.reserve(10));
  CU_TRY(cudaMemsetAsync(h->lat_flag.p, 0, sizeof(int), h->stream));
  const double* pars = h->P ? h->pars.p : nullptr;
  double t = 0.0;
  int* out = h->lat_flag.p;
  void* params[3] = {&pars, &t, &out};
  CU_TRY(cudaLaunchKernel((const void*)h->k_probe, dim3((unsigned)h->sm_count * 8u), dim3(256), params, 0, h->stream));
  int reach = 0;
  CU_TRY(cudaMemcpyAsync(&reach, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  const int fields = (int)((long long)h->N / h->lat_cells_ct);
  if (reach > 16 || 12 * reach > 256 || 2 * reach + 1 > h->lat_cells_ct) return ODEINTR_OK;
  const size_t smem = 2 * (size_t)fields * (256 * (size_t)tile5_cpt(fields) + 12 * (size_t)reach) * sizeof(double);
  // radius 1 and 2 have kernels with the radius folded in (they hand seam / ragged tiles to the general form)
  if ((reach == 1 || reach == 2) && h->k_dopri5_tiled_h[reach] && !std::getenv("ODEINTR_TILE5_GENERAL_ONLY"))
    h->k_dopri5_tiled = h->k_dopri5_tiled_h[reach];
  if (smem > 200 * 1024 ||
      cudaFuncSetAttribute((const void*)h->k_dopri5_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    (void)cudaGetLastError();
    return ODEINTR_OK;
  }
  h->tiled5_halo = reach;
  h->tiled5_smem = smem;
  h->tiled5 = 1;
  // radius <= 2: the interior cells of an attempt go to the warp-autonomous kernel (lattice_dopri5_warp.cuh)
  h->k_warp5 = nullptr; h->warp5_blocks = 0;
  const int wh = reach < 1 ? 1 : reach;
  if (wh <= 2 && h->k_warp5_h[wh] && h->k_warp5_info && !std::getenv("ODEINTR_NO_WARP")) {
    void* wout = h->lat_flag.p;
    void* wparams[1] = {&wout};
    int cpt = 0;
    CU_TRY(cudaLaunchKernel((const void*)h->k_warp5_info, dim3(1), dim3(1), wparams, 0, h->stream));
    CU_TRY(cudaMemcpyAsync(&cpt, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    const size_t wsmem = 8 * (8 * (size_t)fields * 32 * (size_t)cpt * sizeof(double) + 2 * sizeof(unsigned long long));
    int nb = 0;
    if (cpt >= wh && (cpt & 1) && 32 * cpt > 12 * wh && wsmem <= 220 * 1024 &&
        cudaFuncSetAttribute((const void*)h->k_warp5_h[wh], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)h->k_warp5_h[wh], 256, wsmem) == cudaSuccess && nb > 0) {
      h->k_warp5 = h->k_warp5_h[wh];
      h->warp5_cpt = cpt; h->warp5_halo = wh; h->warp5_smem = wsmem; h->warp5_blocks = nb;
    }
    (void)cudaGetLastError();
  }
  return ODEINTR_OK;
}

// ghost cells <- ring neighbours' boundary cells, every field of every buffer, one NCCL group (or local copies for one
// rank)
int lattice_exchange_many(odeintr_system_t h, double* const* bufs, int nb, cudaStream_t stream) {
  const long long g = h->sh_ghost, n = h->sh_n, pitch = h->sh_pitch;
  if (g == 0 || nb == 0) return ODEINTR_OK;
  const int world = h->comm_world, rank = h->comm_rank;
  const long long pad = h->sh_pad;
  // per field: owned cells live at [pad, pad + n); ghost cells at [pad - g, pad) and [pad + n, pad + n + g)
  if (world == 1) {
    for (int b = 0; b < nb; ++b)
      for (int f = 0; f < h->sh_fields; ++f) {
        double* own = bufs[b] + (long long)f * pitch + pad;
        CU_TRY(cudaMemcpyAsync(own - g, own + n - g, g * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        CU_TRY(cudaMemcpyAsync(own + n, own, g * sizeof(double), cudaMemcpyDeviceToDevice, stream));
      }
    return ODEINTR_OK;
  }
  if (!h->comm) return fail(ODEINTR_E_STATE, "odeintr_lattice_comm_init has not been called");
  const int left = (rank + world - 1) % world, right = (rank + 1) % world;
  NCCL_TRY(g_nccl.GroupStart());
  for (int b = 0; b < nb; ++b)
    for (int f = 0; f < h->sh_fields; ++f) {
      double* own = bufs[b] + (long long)f * pitch + pad;
      // order matters when left == right (two ranks): sends (to right, to left) pair with recvs (from left, from right)
      NCCL_TRY(g_nccl.Send(own + n - g, (size_t)g, ncclDouble, right, h->comm, stream));
      NCCL_TRY(g_nccl.Recv(own - g, (size_t)g, ncclDouble, left, h->comm, stream));
      NCCL_TRY(g_nccl.Send(own, (size_t)g, ncclDouble, left, h->comm, stream));
      NCCL_TRY(g_nccl.Recv(own + n, (size_t)g, ncclDouble, right, h->comm, stream));
    }
  NCCL_TRY(g_nccl.GroupEnd());
  return ODEINTR_OK;
}
int lattice_exchange(odeintr_system_t h, double* buf, cudaStream_t stream) {
  double* bufs[1] = {buf};
  return lattice_exchange_many(h, bufs, 1, stream);
}

// ---- slab vectors of a sharded adaptive run: which of them have current ghost cells ------------------------------
bool slab_is_fresh(odeintr_system_t h, const double* p) {
  return std::find(h->slab_fresh.begin(), h->slab_fresh.end(), p) != h->slab_fresh.end();
}
void slab_mark(odeintr_system_t h, const double* p, bool fresh) {
  auto it = std::find(h->slab_fresh.begin(), h->slab_fresh.end(), p);
  if (fresh && it == h->slab_fresh.end()) h->slab_fresh.push_back(p);
  if (!fresh && it != h->slab_fresh.end()) h->slab_fresh.erase(it);
}
// make the ghost cells of (up to two) slab vectors current: one ring exchange for those that are stale
int slab_refresh(odeintr_system_t h, const double* a, const double* b) {
  double* stale[2];
  int n = 0;
  if (a && !slab_is_fresh(h, a)) stale[n++] = const_cast<double*>(a);
  if (b && b != a && !slab_is_fresh(h, b)) stale[n++] = const_cast<double*>(b);
  if (!n) return ODEINTR_OK;
  const int rc = lattice_exchange_many(h, stale, n, h->stream);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) slab_mark(h, stale[i], true);
  return ODEINTR_OK;
}
// max over the ranks of the two words (error bits, reach flag) the attempt kernel left in lat_err
int slab_allreduce_err(odeintr_system_t h) {
  if (h->comm_world <= 1) return ODEINTR_OK;
  if (!h->comm) return fail(ODEINTR_E_STATE, "odeintr_lattice_comm_init has not been called");
  // non-negative doubles order like their bit patterns: the max of the patterns is the pattern of the max
  NCCL_TRY(g_nccl.AllReduce(h->lat_err.p, h->lat_err.p, 2, ncclUint64, ncclMax, h->comm, h->stream));
  return ODEINTR_OK;
}

#include "lattice_adaptive.h"

// Fixed-step Cash-Karp (method "rk54") on a large system: six combo launches, the result replaces the state.
int lattice_ck54_fixed_step(odeintr_system_t h, double t, double dt) {
  const size_t N = (size_t)h->N;
  CU_TRY(h->lat_work.reserve(9 * N));
  double* w = h->lat_work.p;
  if (!h->lat_work_primed) {
    // state-like vectors start as copies of x, derivative-like ones as zeros (entries no loop writes)
    for (int j = 0; j < 3; ++j)
      CU_TRY(cudaMemcpyAsync(w + (size_t)j * N, h->lat_x, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemsetAsync(w + 3 * N, 0, 6 * N * sizeof(double), h->stream));
    h->lat_work_primed = true;
  }
  lattice_ck54 ck;
  ck.core.h = h;
  ck.core.xtA = w; ck.core.xtB = w + N;
  ck.k1 = w + 3 * N;
  ck.core.k2 = w + 4 * N; ck.core.k3 = w + 5 * N; ck.core.k4 = w + 6 * N; ck.core.k5 = w + 7 * N; ck.core.k6 = w + 8 * N;
  ck.stages(h->lat_x, t, dt, w + 2 * N, false);
  if (ck.core.rc) return ck.core.rc;
  CU_TRY(cudaMemcpyAsync(h->lat_x, w + 2 * N, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return ODEINTR_OK;
}

// Fixed-step Fehlberg 7(8) (method "rk78") on a large system: thirteen combo launches, the result replaces the state.
int lattice_rkf78_fixed_step(odeintr_system_t h, double t, double dt) {
  const size_t N = (size_t)h->N;
  CU_TRY(h->lat_work.reserve(16 * N));
  double* w = h->lat_work.p;
  if (!h->lat_work_primed) {
    for (int j = 0; j < 3; ++j)
      CU_TRY(cudaMemcpyAsync(w + (size_t)j * N, h->lat_x, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemsetAsync(w + 3 * N, 0, 13 * N * sizeof(double), h->stream));
    h->lat_work_primed = true;
  }
  lattice_rkf78 rk;
  rk.core.h = h;
  rk.xt[0] = w; rk.xt[1] = w + N;
  for (int j = 0; j < 13; ++j) rk.k[j] = w + (size_t)(3 + j) * N;
  rk.stages(h->lat_x, t, dt, w + 2 * N, false);
  if (rk.core.rc) return rk.core.rc;
  CU_TRY(cudaMemcpyAsync(h->lat_x, w + 2 * N, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return ODEINTR_OK;
}

// dopri5 as a plain fixed-step stepper (method "rk5") on a large system.  explicit_error_stepper_fsal_base::do_step
// evaluates f(x) on the first call only and hands k7 of every step to the next one as k1; a step is one launch of
// the temporally blocked attempt kernel (or the seven combo launches), the error estimate is ignored.
int lattice_rk5_fixed_step(odeintr_system_t h, double t, double dt) {
  const size_t N = (size_t)h->N;
  int rc;
  CU_TRY(h->lat_work.reserve(11 * N));
  CU_TRY(h->lat_err.reserve(2));
  if ((rc = tiled5_decide(h))) return rc;
  double* w = h->lat_work.p;
  lattice_dopri5 core;
  core.h = h;
  core.xtA = w; core.xtB = w + N;
  double* xnew = w + 2 * N;
  core.k2 = w + 3 * N; core.k3 = w + 4 * N; core.k4 = w + 5 * N; core.k5 = w + 6 * N; core.k6 = w + 7 * N;
  double* k1 = w + (size_t)(8 + (h->lat_rk5_parity & 1)) * N;
  double* k7 = w + (size_t)(8 + ((h->lat_rk5_parity + 1) & 1)) * N;
  if (!h->lat_work_primed) {
    for (int j = 0; j < 3; ++j)
      CU_TRY(cudaMemcpyAsync(w + (size_t)j * N, h->lat_x, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemsetAsync(w + 3 * N, 0, 8 * N * sizeof(double), h->stream));
    h->lat_work_primed = true;
    core.rhs(h->lat_x, k1, t);  // first call of this integrate_* invocation
  }
  core.step_with_error(h->lat_x, k1, t, xnew, k7, dt);
  if (core.rc) return core.rc;
  CU_TRY(cudaMemcpyAsync(h->lat_x, xnew, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->lat_rk5_parity ^= 1;
  return ODEINTR_OK;
}

// rk4 / euler of a GENERAL large-system snippet (odeintr_lattice_serial_rhs): Boost's own unfused schedule -- every
// stage derivative is a stored vector, the combinations are odeintr_lattice_combo launches in scale_sumN's order
// (coefficient * dt first, left to right, zero tableau entries skipped like everywhere else).
int lattice_serial_fixed_step(odeintr_system_t h, double t, double dt) {
  const size_t N = (size_t)h->N;
  CU_TRY(h->lat_work.reserve(7 * N));
  CU_TRY(h->lat_err.reserve(2));
  double* w = h->lat_work.p;
  if (!h->lat_work_primed) {
    for (int j = 0; j < 3; ++j)
      CU_TRY(cudaMemcpyAsync(w + (size_t)j * N, h->lat_x, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemsetAsync(w + 3 * N, 0, 4 * N * sizeof(double), h->stream));
    h->lat_work_primed = true;
  }
  lattice_dopri5 core;
  core.h = h;
  double *xtA = w, *xtB = w + N, *xnew = w + 2 * N, *k1 = w + 3 * N, *k2 = w + 4 * N, *k3 = w + 5 * N, *k4 = w + 6 * N;
  const double* x = h->lat_x;
  if (h->lattice_stages == 1) {  // euler: x + dt * f(x)
    core.combo(x, x, 0, nullptr, nullptr, dt, xnew, k1, t);
  } else {
    const double half = 1.0 / 2.0, b1 = 1.0 / 6.0, b2 = 1.0 / 3.0;
    core.combo(x, x, 0, nullptr, nullptr, half * dt, xtA, k1, t);
    core.combo(xtA, x, 0, nullptr, nullptr, half * dt, xtB, k2, t + half * dt);
    core.combo(xtB, x, 0, nullptr, nullptr, 1.0 * dt, xtA, k3, t + half * dt);
    const double* v[] = {k1, k2, k3};
    const double c[] = {b1 * dt, b2 * dt, b2 * dt};
    core.combo(xtA, x, 3, v, c, b1 * dt, xnew, k4, t + 1.0 * dt);
  }
  if (core.rc) return core.rc;
  CU_TRY(cudaMemcpyAsync(h->lat_x, xnew, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return ODEINTR_OK;
}

int run_adaptive(odeintr_system_t h, int mode, double t0, double t1, double dt, const double* times, size_t n_times,
                 bool record, bool wide);
// Large system under an adaptive stepper: the shared integrate loops run on the host over device vectors.
int run_lattice_adaptive(odeintr_system_t h, int mode, double t0, double t1, double dt, const double* times,
                         size_t n_times, bool record) {
  if (!h->k_combo) return fail(ODEINTR_E_STATE, "system was not compiled for adaptive stepping of a large system");
  if (h->m > 1 && !h->sharded) {
    // an ensemble of large systems: one warp / block per trajectory, every trajectory with its own step sizes
    return run_adaptive(h, mode, mode == ODEINTR_MODE_TIMES ? times[0] : t0, mode == ODEINTR_MODE_TIMES ? times[n_times - 1] : t1,
                        dt, times, n_times, record, true);
  }
  if (h->P && h->par_sets != 1) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  // One GPU: the state is h->x (N doubles).  After odeintr_lattice_shard the same loops run on this rank's slab of the
  // lattice -- every rank makes the same calls with the same arguments (collective): the error norm of an attempt is
  // the maximum over all slabs, so all ranks take the same decisions and the same steps.
  const bool slab = h->sharded;
  if (slab) {
    if (h->lattice_method != 1 && h->lattice_method != 2)
      return fail(ODEINTR_E_STATE, "sharded adaptive stepping is provided for rk5_i and rk5_a");
    if (!h->k_lattice_sharded) return fail(ODEINTR_E_STATE, "the generated translation unit has no sharded stage kernel");
  } else if (h->m != 1) {
    return fail(ODEINTR_E_STATE, "Invalid initial state");
  }
  const size_t N = slab ? (size_t)h->sh_fields * (size_t)h->sh_pitch : (size_t)h->N;
  double* const xstate = slab ? h->sh_x : h->x.p;
  h->lat_vec_len = N;
  h->slab_fresh.clear();
  const bool bs_family = h->lattice_method == 5 || h->lattice_method == 6;
  if (bs_family) {  // Bulirsch-Stoer: work vectors from a pool that grows with the orders the controller reaches
    if (h->lat_pool_len != N) {
      for (double* v : h->lat_pool) cudaFree(v);
      h->lat_pool.clear();
      h->lat_pool_len = N;
    }
    h->lat_pool_used = 0;
  }
  CU_TRY(h->lat_work.reserve((bs_family ? 0 : h->lattice_method == 4 ? 18 : 13) * N + 1));
  CU_TRY(h->lat_err.reserve(2));  // max-norm error bits, reach flag
  {
    const int rc5 = tiled5_decide(h);
    if (rc5) return rc5;
  }
  if (slab) {
    if (h->tiled5 != 1)
      return fail(ODEINTR_E_STATE, "sharded adaptive stepping needs the temporally blocked dopri5 kernel (stencil radius <= 16)");
    if (h->sh_ghost < 6 * h->tiled5_halo)
      return fail(ODEINTR_E_INVALID, "ghost zone narrower than 6 x the stencil radius (" + std::to_string(h->tiled5_halo) +
                                         "): declare the radius to odeintr_lattice_shard");
    CU_TRY(h->lat_negzero.reserve(N));
    CU_TRY(cudaMemsetAsync(h->lat_negzero.p, 0, N * sizeof(double), h->stream));
    odeintr_b200::launch_fill_negzero(h->lat_negzero.p, (long long)N, h->stream);
    g_total_launches += 1;
  }
  double* w = h->lat_work.p;
  // state-like vectors start as copies of x, derivative-like ones as zeros: entries no user loop writes keep
  // their value / a zero derivative through every stage
  double* state_like[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  double* deriv_like[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (!bs_family) {
    for (int j = 0; j < 5; ++j) {
      state_like[j] = w + (size_t)j * N;
      CU_TRY(cudaMemcpyAsync(state_like[j], xstate, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    CU_TRY(cudaMemsetAsync(w + 5 * N, 0, 8 * N * sizeof(double), h->stream));
    for (int j = 0; j < 8; ++j) deriv_like[j] = w + (size_t)(5 + j) * N;
  }

  lattice_dopri5 core;
  core.h = h;
  core.slab = slab;
  core.xtA = state_like[0]; core.xtB = state_like[1];
  core.k2 = deriv_like[0]; core.k3 = deriv_like[1]; core.k4 = deriv_like[2]; core.k5 = deriv_like[3];
  core.k6 = deriv_like[4];
  std::vector<double> rec_t;
  size_t cap_rows = 0;
  if (record) {
    size_t guess = mode == ODEINTR_MODE_TIMES ? n_times : 16;
    if (mode == ODEINTR_MODE_CONST) guess = (size_t)const_step_count(t0, t1, dt) + 2;
    cudaError_t ce = h->out.reserve(guess * N);
    if (ce != cudaSuccess) return fail(ODEINTR_E_CUDA, std::string("cannot allocate observer record: ") + cudaGetErrorString(ce));
    cap_rows = h->out.cap / N;
  }
  lattice_observer obs{h, &rec_t, cap_rows};
  odeintr_b200::null_observer nobs;
  lat_sys sys;
  lat_state x(h, xstate);
  int rc;
  if ((rc = begin_timing(h))) return rc;
  bool ok = true;  // failures are carried by the counters' status word
  lat_counters cnt;
  if (h->lattice_method == 5) {
    lattice_bs_controlled st;
    st.core = core;
    st.cnt.budget = h->max_tries;
    st.init();
    if (st.core.rc) return st.core.rc;
    double dtc = dt, tc = t0;
    if (mode == ODEINTR_MODE_TIMES)
      ok = record ? odeintr_b200::integrate_times_controlled(st, sys, x, times, (long long)n_times, dtc, obs)
                  : odeintr_b200::integrate_times_controlled(st, sys, x, times, (long long)n_times, dtc, nobs);
    else if (mode == ODEINTR_MODE_CONST)
      ok = record ? odeintr_b200::integrate_const_controlled(st, sys, x, t0, t1, dtc, obs)
                  : odeintr_b200::integrate_const_controlled(st, sys, x, t0, t1, dtc, nobs);
    else
      ok = record ? odeintr_b200::integrate_adaptive_controlled(st, sys, x, tc, t1, dtc, obs)
                  : odeintr_b200::integrate_adaptive_controlled(st, sys, x, tc, t1, dtc, nobs);
    cnt = st.cnt;
    if (st.core.rc) return st.core.rc;
  } else if (h->lattice_method == 6) {
    lattice_bsd_dense st;
    st.core = core;
    st.cnt.budget = h->max_tries;
    st.init();
    if (st.core.rc) return st.core.rc;
    if (mode == ODEINTR_MODE_TIMES)
      ok = record ? odeintr_b200::integrate_times_dense(st, sys, x, times, (long long)n_times, dt, obs)
                  : odeintr_b200::integrate_times_dense(st, sys, x, times, (long long)n_times, dt, nobs);
    else if (mode == ODEINTR_MODE_CONST)
      ok = record ? odeintr_b200::integrate_const_dense(st, sys, x, t0, t1, dt, obs)
                  : odeintr_b200::integrate_const_dense(st, sys, x, t0, t1, dt, nobs);
    else
      ok = record ? odeintr_b200::integrate_adaptive_dense(st, sys, x, t0, t1, dt, obs)
                  : odeintr_b200::integrate_adaptive_dense(st, sys, x, t0, t1, dt, nobs);
    cnt = st.cnt;
    if (st.core.rc) return st.core.rc;
  } else if (h->lattice_method == 2) {
    lattice_dopri5_dense st;
    st.ctl.core = core;
    st.ctl.cnt.budget = h->max_tries;
    st.x_cur.h = h; st.x_cur.p = state_like[2];  // (assignment of lat_state copies values, not handles)
    st.x_old.h = h; st.x_old.p = state_like[3];
    st.d_cur = deriv_like[5]; st.d_old = deriv_like[6];
    if (mode == ODEINTR_MODE_TIMES)
      ok = record ? odeintr_b200::integrate_times_dense(st, sys, x, times, (long long)n_times, dt, obs)
                  : odeintr_b200::integrate_times_dense(st, sys, x, times, (long long)n_times, dt, nobs);
    else if (mode == ODEINTR_MODE_CONST)
      ok = record ? odeintr_b200::integrate_const_dense(st, sys, x, t0, t1, dt, obs)
                  : odeintr_b200::integrate_const_dense(st, sys, x, t0, t1, dt, nobs);
    else
      ok = record ? odeintr_b200::integrate_adaptive_dense(st, sys, x, t0, t1, dt, obs)
                  : odeintr_b200::integrate_adaptive_dense(st, sys, x, t0, t1, dt, nobs);
    cnt = st.ctl.cnt;
    if (st.ctl.core.rc) return st.ctl.core.rc;
  } else if (h->lattice_method == 4) {
    lattice_rkf78_controlled st;
    st.rk.core = core;
    st.cnt.budget = h->max_tries;
    st.rk.xt[0] = state_like[0]; st.rk.xt[1] = state_like[1];
    st.xnew = state_like[2];
    for (int j = 0; j < 8; ++j) st.rk.k[j] = deriv_like[j];
    CU_TRY(cudaMemsetAsync(w + 13 * N, 0, 5 * N * sizeof(double), h->stream));
    for (int j = 8; j < 13; ++j) st.rk.k[j] = w + (size_t)(5 + j) * N;
    double dtc = dt, tc = t0;
    if (mode == ODEINTR_MODE_TIMES)
      ok = record ? odeintr_b200::integrate_times_controlled(st, sys, x, times, (long long)n_times, dtc, obs)
                  : odeintr_b200::integrate_times_controlled(st, sys, x, times, (long long)n_times, dtc, nobs);
    else if (mode == ODEINTR_MODE_CONST)
      ok = record ? odeintr_b200::integrate_const_controlled(st, sys, x, t0, t1, dtc, obs)
                  : odeintr_b200::integrate_const_controlled(st, sys, x, t0, t1, dtc, nobs);
    else
      ok = record ? odeintr_b200::integrate_adaptive_controlled(st, sys, x, tc, t1, dtc, obs)
                  : odeintr_b200::integrate_adaptive_controlled(st, sys, x, tc, t1, dtc, nobs);
    cnt = st.cnt;
    if (st.rk.core.rc) return st.rk.core.rc;
  } else if (h->lattice_method == 3) {
    lattice_ck54_controlled st;
    st.core = core;
    st.cnt.budget = h->max_tries;
    st.k1 = deriv_like[5]; st.xnew = state_like[2];
    double dtc = dt, tc = t0;
    if (mode == ODEINTR_MODE_TIMES)
      ok = record ? odeintr_b200::integrate_times_controlled(st, sys, x, times, (long long)n_times, dtc, obs)
                  : odeintr_b200::integrate_times_controlled(st, sys, x, times, (long long)n_times, dtc, nobs);
    else if (mode == ODEINTR_MODE_CONST)
      ok = record ? odeintr_b200::integrate_const_controlled(st, sys, x, t0, t1, dtc, obs)
                  : odeintr_b200::integrate_const_controlled(st, sys, x, t0, t1, dtc, nobs);
    else
      ok = record ? odeintr_b200::integrate_adaptive_controlled(st, sys, x, tc, t1, dtc, obs)
                  : odeintr_b200::integrate_adaptive_controlled(st, sys, x, tc, t1, dtc, nobs);
    cnt = st.cnt;
    if (st.core.rc) return st.core.rc;
  } else {
    lattice_dopri5_controlled st;
    st.core = core;
    st.cnt.budget = h->max_tries;
    st.dxdt = deriv_like[5]; st.dnew = deriv_like[6]; st.xnew = state_like[2];
    double dtc = dt, tc = t0;
    if (mode == ODEINTR_MODE_TIMES)
      ok = record ? odeintr_b200::integrate_times_controlled(st, sys, x, times, (long long)n_times, dtc, obs)
                  : odeintr_b200::integrate_times_controlled(st, sys, x, times, (long long)n_times, dtc, nobs);
    else if (mode == ODEINTR_MODE_CONST)
      ok = record ? odeintr_b200::integrate_const_controlled(st, sys, x, t0, t1, dtc, obs)
                  : odeintr_b200::integrate_const_controlled(st, sys, x, t0, t1, dtc, nobs);
    else
      ok = record ? odeintr_b200::integrate_adaptive_controlled(st, sys, x, tc, t1, dtc, obs)
                  : odeintr_b200::integrate_adaptive_controlled(st, sys, x, tc, t1, dtc, nobs);
    cnt = st.cnt;
    if (st.core.rc) return st.core.rc;
    if (x.p != xstate) {  // accepted steps swap the state vector with the stepper's: bring the result home
      CU_TRY(cudaMemcpyAsync(xstate, x.p, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      if (slab) slab_mark(h, xstate, slab_is_fresh(h, x.p));
    }
  }
  if (!ok && cnt.status == ODEINTR_STATUS_OK) cnt.status = ODEINTR_STATUS_NONFINITE;
  if ((rc = end_timing(h))) return rc;
  if (obs.rc) return fail(ODEINTR_E_CUDA, "cannot grow the observer record");
  CU_TRY(h->accepted.reserve(1)); CU_TRY(h->rejected.reserve(1)); CU_TRY(h->status.reserve(1));
  CU_TRY(cudaMemcpyAsync(h->accepted.p, &cnt.accepted, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemcpyAsync(h->rejected.p, &cnt.rejected, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemcpyAsync(h->status.p, &cnt.status, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->last_steps = cnt.accepted;
  if (record) {
    h->rec_t = rec_t;
    h->rows = rec_t.size();
    h->out_m = 1;
    h->ragged = false;
  }
  if (cnt.status & ODEINTR_STATUS_STEP_OVERFLOW)
    return fail(ODEINTR_E_STEP, "Max number of iterations exceeded (500). A new step size was not found.");
  if (cnt.status & ODEINTR_STATUS_STEP_BUDGET)
    return fail(ODEINTR_E_STEP, "step budget exceeded: more than " + std::to_string(h->max_tries) +
                                    " step attempts in one call (see odeintr_set_max_tries)");
  if (cnt.status & ODEINTR_STATUS_NONFINITE) return fail(ODEINTR_E_STEP, "integration stopped making progress (non-finite time)");
  return ODEINTR_OK;
}

// Time column integrate_const records with a dense-output stepper: instance independent.
std::vector<double> dense_const_times(double t0, double t1, double dt, bool* optional_last = nullptr) {
  if (optional_last) *optional_last = false;
  std::vector<double> rec;
  double time = t0;
  rec.push_back(time);
  time += dt;
  long long k = 1;
  while (less_eq_with_sign(time + dt, t1, dt)) {
    rec.push_back(time);
    ++k;
    time = t0 + static_cast<double>(k) * dt;
  }
  if (less_eq_with_sign(time, t1, dt)) {
    rec.push_back(time);
    // One more sample can appear: when the stepper's last real step (the one landing on t1) is taken while two or
    // more samples are still ahead, Boost's inner emission loop also delivers the sample after this one, provided
    // it is not beyond t1.  Whether that happens depends on the step sizes, i.e. on the instance.
    ++k;
    const double extra = t0 + static_cast<double>(k) * dt;
    if (optional_last && less_eq_with_sign(extra, t1, dt)) { rec.push_back(extra); *optional_last = true; }
  }
  return rec;
}

// Controlled / dense-output steppers: one launch, per-instance step control inside the kernel.
int run_adaptive(odeintr_system_t h, int mode, double t0, double t1, double dt, const double* times, size_t n_times,
                 bool record, bool wide) {
  // wide: an ensemble of compile_sys_openmp systems under rk5_i / rk5_a, one warp / block per trajectory
  cudaKernel_t kern = wide ? (h->lattice_method >= 1 && h->lattice_method <= 6 ? h->k_lat_ens_a[h->lattice_method] : nullptr) : h->k_adaptive;
  if (!kern)
    return fail(ODEINTR_E_STATE, wide ? "adaptive ensembles of large systems are provided on 1-D lattices of 3 ... 2048 cells with at "
                                        "most 4 states per thread (fields x cells / threads); run the trajectories one at a time"
                                      : "system was not compiled with an adaptive stepper");
  if (h->m == 0) return fail(ODEINTR_E_STATE, "Invalid initial state");
  const size_t m = h->m, N = (size_t)h->N;
  odeintr_lattice_ensemble_adaptive_args wa;
  std::memset(&wa, 0, sizeof(wa));
  odeintr_adaptive_args& a = wa.p;
  int rc = param_view(h, a.pars, a.par_ld, a.par_inc);
  if (rc) return rc;
  unsigned wide_block = 0, wide_grid = 0;
  size_t wide_smem = 0;
  if (wide) {
    // stencil radius of the loop body, measured with the system's parameters in place (odeintr_lattice_probe)
    if (!h->k_probe || h->lat_cells_ct <= 0) return fail(ODEINTR_E_STATE, "the generated translation unit has no probe kernel");
    if (h->P && !h->pars.p) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
    int reach = 0;
    {
      CU_TRY(h->lat_flag.reserve(10));
      CU_TRY(cudaMemsetAsync(h->lat_flag.p, 0, sizeof(int), h->stream));
      const double* ppars = h->P ? h->pars.p : nullptr;
      double pt = 0.0;
      int* pout = h->lat_flag.p;
      void* pparams[3] = {&ppars, &pt, &pout};
      CU_TRY(cudaLaunchKernel((const void*)h->k_probe, dim3((unsigned)h->sm_count * 8u), dim3(256), pparams, 0, h->stream));
      CU_TRY(cudaMemcpyAsync(&reach, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CU_TRY(cudaStreamSynchronize(h->stream));
    }
    if (reach > 32) return fail(ODEINTR_E_INVALID, "ensembles of large systems need a stencil radius of at most 32 cells");
    const long long L = h->lat_cells_ct, halo = reach < 1 ? 1 : reach;
    const int fields = (int)((long long)h->N / L);
    if (2 * halo + 1 > L) return fail(ODEINTR_E_INVALID, "stencil radius too wide for the lattice");
    // threads per trajectory and trajectories per block, as lattice_ensemble.cuh derives them from the lattice length
    const unsigned nt = (unsigned)(L > 2048 ? 512 : (L >= 256 ? 256 : (L + 31) / 32 * 32));
    const unsigned group = nt == 32 ? 4 : 1;
    wide_smem = group * (2 * (size_t)fields * (size_t)(L + 2 * halo) + 2 * (nt / 32)) * sizeof(double);
    if (wide_smem > 200 * 1024) return fail(ODEINTR_E_INVALID, "lattice too large for one block's shared memory");
    CU_TRY(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem));
    wide_block = nt * group;
    wide_grid = (unsigned)((m + group - 1) / group);
    CU_TRY(h->lat_flag.reserve(10));
    CU_TRY(cudaMemsetAsync(h->lat_flag.p, 0, sizeof(int), h->stream));
    wa.halo = halo;
    wa.flag = h->lat_flag.p;
  }
  auto run_kernel = [&](odeintr_adaptive_args& args, unsigned grid, unsigned block) -> int {
    if (!wide) return launch(h, kern, h->stream, grid, block, &args);
    odeintr_lattice_ensemble_adaptive_args w2 = wa;
    w2.p = args;
    void* params[1] = {&w2};
    CU_TRY(cudaLaunchKernel((const void*)kern, dim3(wide_grid), dim3(wide_block), params, wide_smem, h->stream));
    h->last_launches += 1;
    g_total_launches += 1;
    return ODEINTR_OK;
  };
  CU_TRY(h->accepted.reserve(m)); CU_TRY(h->rejected.reserve(m)); CU_TRY(h->status.reserve(m));
  CU_TRY(h->rows_inst.reserve(m)); CU_TRY(h->fail_flags.reserve(1));
  std::vector<double> rec_t;
  bool optional_last = false;  // rec_t ends with a sample that only some instances produce
  if (mode == ODEINTR_MODE_TIMES) {
    CU_TRY(h->times_dev.reserve(n_times));
    if ((rc = upload(h, h->times_dev.p, times, n_times * sizeof(double)))) return rc;
    a.times = h->times_dev.p;
    a.n_times = (long long)n_times;
    rec_t.assign(times, times + n_times);
  } else if (mode == ODEINTR_MODE_CONST && record) {
    if (h->dense) rec_t = dense_const_times(t0, t1, dt, &optional_last);
    else {
      const long long n = const_step_count(t0, t1, dt);
      rec_t.resize((size_t)n + 1);
      for (long long s = 0; s <= n; ++s) rec_t[(size_t)s] = t0 + static_cast<double>(s) * dt;
    }
  }
  a.x = h->x.p;
  a.m = (long long)m; a.ld = (long long)m;
  a.t0 = t0; a.t1 = t1; a.dt = dt;
  a.atol = h->atol; a.rtol = h->rtol;
  a.mode = mode;
  a.max_tries = h->max_tries;
  a.accepted = h->accepted.p; a.rejected = h->rejected.p; a.status = h->status.p; a.rows = h->rows_inst.p;
  a.fail_flags = h->fail_flags.p;
  CU_TRY(h->rows_min.reserve(1));
  if (optional_last) {
    const int big = 0x7fffffff;
    CU_TRY(cudaMemcpyAsync(h->rows_min.p, &big, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    a.rows_min = h->rows_min.p;
  }
  const unsigned block = (unsigned)h->block_adaptive;
  const unsigned grid = (unsigned)((m + block - 1) / block);
  CU_TRY(cudaMemsetAsync(h->fail_flags.p, 0, sizeof(int), h->stream));
  if ((rc = begin_timing(h))) return rc;
  // K6: instance order.  The pilot and the sort are part of the call (and of its device time).
  if (!wide && h->order_mode == 1 && h->k_pilot && m >= 64 && m < (1ull << 31)) {
    CU_TRY(h->order_keys.reserve(m)); CU_TRY(h->order_keys_sorted.reserve(m));
    CU_TRY(h->perm.reserve(m)); CU_TRY(h->perm_iota.reserve(m));
    const size_t tb = odeintr_b200::sort_pairs_temp_bytes((long long)m);
    CU_TRY(h->sort_temp.reserve(tb));
    unsigned* keys = h->order_keys.p;
    int attempts = h->pilot_attempts;
    void* pparams[3] = {&a, &keys, &attempts};
    CU_TRY(cudaLaunchKernel((const void*)h->k_pilot, dim3(grid), dim3(block), pparams, 0, h->stream));
    h->last_launches += 1;
    g_total_launches += 1;
    if (odeintr_b200::launch_sort_by_key(h->sort_temp.p, tb, h->order_keys.p, h->order_keys_sorted.p, h->perm_iota.p,
                                         h->perm.p, (long long)m, h->stream) != 0)
      return fail(ODEINTR_E_CUDA, "instance-order sort failed");
    g_total_launches += 2;
    a.perm = h->perm.p;
  } else if (!wide && h->order_mode == 2 && h->perm_m == m) {
    a.perm = h->perm.p;
  }
  bool ragged = false;
  int max_rows = 0;
  if (mode == ODEINTR_MODE_ADAPTIVE && record) {
    // Ragged record (one row per accepted step): a dry run on a copy of the state counts the rows of every
    // instance; the arithmetic is deterministic, so the recording run below takes exactly the same steps.
    CU_TRY(h->stage.reserve(N * m));
    CU_TRY(cudaMemcpyAsync(h->stage.p, h->x.p, N * m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    odeintr_adaptive_args dry = a;
    dry.x = h->stage.p;
    if ((rc = run_kernel(dry, grid, block))) return rc;
    std::vector<int> rows_host(m);
    CU_TRY(cudaMemcpyAsync(rows_host.data(), h->rows_inst.p, m * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < m; ++i) max_rows = std::max(max_rows, rows_host[i]);
    if ((rc = reserve_out(h, (size_t)max_rows))) return rc;
    CU_TRY(h->t_rec.reserve((size_t)max_rows * m));
    a.out = h->out.p; a.t_rec = h->t_rec.p; a.max_rows = max_rows;
    CU_TRY(cudaMemsetAsync(h->fail_flags.p, 0, sizeof(int), h->stream));
    ragged = true;
  } else if (record) {
    if ((rc = reserve_out(h, rec_t.size()))) return rc;
    a.out = h->out.p;
    a.max_rows = (int)rec_t.size();
  }
  if ((rc = run_kernel(a, grid, block))) return rc;
  if ((rc = end_timing(h))) return rc;
  if (wide) {
    int reach_flag = 0;
    CU_TRY(cudaMemcpyAsync(&reach_flag, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    if (reach_flag)
      return fail(ODEINTR_E_INVALID, "the system definition reads x further than the stencil radius (" +
                                         std::to_string(wa.halo) + " cells) from the cell it computes");
  }
  int flags = 0, rows_min = 0x7fffffff;
  CU_TRY(cudaMemcpyAsync(&flags, h->fail_flags.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (optional_last) CU_TRY(cudaMemcpyAsync(&rows_min, h->rows_min.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (optional_last && !flags && rows_min < (int)rec_t.size()) rec_t.resize((size_t)std::max(rows_min, 1));
  h->last_steps = 0;
  if (record) {
    h->rec_t = rec_t;
    h->rows = ragged ? (size_t)max_rows : rec_t.size();
    h->out_m = m;
    h->ragged = ragged;
    h->max_rows = max_rows;
  }
  if (flags & ODEINTR_STATUS_STEP_OVERFLOW)
    return fail(ODEINTR_E_STEP, "Max number of iterations exceeded (500). A new step size was not found.");
  if (flags & ODEINTR_STATUS_STEP_BUDGET)
    return fail(ODEINTR_E_STEP, "step budget exceeded: an instance needed more than " + std::to_string(h->max_tries) +
                                    " step attempts in one call (see odeintr_set_max_tries)");
  return ODEINTR_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* odeintr_last_error(void) { return g_err.c_str(); }

int odeintr_compile_cubin(const char* cuda_src, unsigned flags, void* cubin, size_t cubin_cap, size_t* cubin_size) {
  if (!cuda_src) return fail(ODEINTR_E_INVALID, "null source");
  std::vector<char> image;
  std::string log;
  int rc = nvrtc_to_cubin(cuda_src, flags, image, log);
  if (rc) return rc;
  if (cubin_size) *cubin_size = image.size();
  if (cubin && cubin_cap >= image.size()) std::memcpy(cubin, image.data(), image.size());
  return ODEINTR_OK;
}

int odeintr_compile(const char* cuda_src, const char* name, int sys_dim, int n_pars, double atol, double rtol,
                    unsigned flags, odeintr_system_t* out) {
  if (!cuda_src || !out) return fail(ODEINTR_E_INVALID, "null argument");
  if (sys_dim < 1) return fail(ODEINTR_E_INVALID, "system dimension must be >= 1");
  if (n_pars < 0) return fail(ODEINTR_E_INVALID, "negative parameter count");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(ODEINTR_E_CUDA, std::string("no CUDA device available (") +
                                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                    "); odeintr_b200 has no CPU fallback");
  odeintr_system_s* h = new odeintr_system_s();
  h->name = name ? name : "sys";
  h->N = sys_dim; h->P = n_pars; h->atol = atol; h->rtol = rtol; h->flags = flags;
  h->dense = (flags & ODEINTR_FLAG_DENSE_OUTPUT) != 0;
  h->lattice_stages = (flags & ODEINTR_FLAG_LATTICE_EULER) ? 1 : ((flags & ODEINTR_FLAG_LATTICE_CK54) ? 6 :
                      ((flags & ODEINTR_FLAG_LATTICE_RKF78) ? 13 : ((flags & ODEINTR_FLAG_LATTICE_RK5) ? 7 : 4)));
  h->lattice_method = (flags & ODEINTR_FLAG_LATTICE_ADAPTIVE) ? ((flags & ODEINTR_FLAG_DENSE_OUTPUT) ? 2 : 1) : 0;
  if ((flags & ODEINTR_FLAG_LATTICE_CK54) && h->lattice_method) h->lattice_method = 3;
  if ((flags & ODEINTR_FLAG_LATTICE_RKF78) && h->lattice_method) h->lattice_method = 4;
  if ((flags & ODEINTR_FLAG_LATTICE_BS) && h->lattice_method) h->lattice_method = (flags & ODEINTR_FLAG_DENSE_OUTPUT) ? 6 : 5;
  std::vector<char> image;
  int rc = nvrtc_to_cubin(cuda_src, flags, image, h->log);
  if (rc) { delete h; return rc; }
  auto bail = [&](const std::string& what, cudaError_t ce) {
    std::string msg = what + ": " + cudaGetErrorString(ce);
    odeintr_destroy(h);
    return fail(ODEINTR_E_CUDA, msg);
  };
  if ((e = cudaGetDevice(&h->device)) != cudaSuccess) return bail("cudaGetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  h->sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    odeintr_destroy(h);
    return fail(ODEINTR_E_CUDA, std::string("device '") + prop.name + "' is not sm_100-class; this engine is B200-only");
  }
  if ((e = cudaLibraryLoadData(&h->lib, image.data(), nullptr, nullptr, 0, nullptr, nullptr, 0)) != cudaSuccess)
    return bail("cudaLibraryLoadData", e);
  // whichever stepper kernels the generated TU instantiated
  if (cudaLibraryGetKernel(&h->k_plain, h->lib, "odeintr_plain_ensemble") != cudaSuccess) h->k_plain = nullptr;
  if (cudaLibraryGetKernel(&h->k_adaptive, h->lib, "odeintr_adaptive_ensemble") != cudaSuccess) h->k_adaptive = nullptr;
  if (cudaLibraryGetKernel(&h->k_pilot, h->lib, "odeintr_adaptive_pilot") != cudaSuccess) h->k_pilot = nullptr;
  if (cudaLibraryGetKernel(&h->k_lattice, h->lib, "odeintr_lattice_stage") != cudaSuccess) h->k_lattice = nullptr;
  if (cudaLibraryGetKernel(&h->k_lattice_sharded, h->lib, "odeintr_lattice_stage_sharded") != cudaSuccess)
    h->k_lattice_sharded = nullptr;
  if (cudaLibraryGetKernel(&h->k_combo, h->lib, "odeintr_lattice_combo") != cudaSuccess) h->k_combo = nullptr;
  if (cudaLibraryGetKernel(&h->k_tiled, h->lib, "odeintr_lattice_rk4_tiled") != cudaSuccess) h->k_tiled = nullptr;
  if (cudaLibraryGetKernel(&h->k_tiled_int, h->lib, "odeintr_lattice_rk4_tiled_interior") != cudaSuccess)
    h->k_tiled_int = nullptr;
  if (cudaLibraryGetKernel(&h->k_tiled_tma, h->lib, "odeintr_lattice_rk4_tiled_tma") != cudaSuccess)
    h->k_tiled_tma = nullptr;
  if (cudaLibraryGetKernel(&h->k_warp_h[1], h->lib, "odeintr_lattice_rk4_warp_h1") != cudaSuccess) h->k_warp_h[1] = nullptr;
  if (cudaLibraryGetKernel(&h->k_warp_h[2], h->lib, "odeintr_lattice_rk4_warp_h2") != cudaSuccess) h->k_warp_h[2] = nullptr;
  if (cudaLibraryGetKernel(&h->k_warp_info, h->lib, "odeintr_lattice_warp_info") != cudaSuccess) h->k_warp_info = nullptr;
  if (cudaLibraryGetKernel(&h->k_lat_ens, h->lib, "odeintr_lattice_ensemble") != cudaSuccess) h->k_lat_ens = nullptr;
  if (cudaLibraryGetKernel(&h->k_lat_ens_plain, h->lib, "odeintr_lattice_ensemble_plain") != cudaSuccess) h->k_lat_ens_plain = nullptr;
  {
    static const char* const names[7] = {nullptr, "odeintr_lattice_ensemble_dopri5_controlled", "odeintr_lattice_ensemble_dopri5_dense",
                                         "odeintr_lattice_ensemble_rk54_controlled", "odeintr_lattice_ensemble_rk78_controlled",
                                         "odeintr_lattice_ensemble_bs_controlled", "odeintr_lattice_ensemble_bsd_dense"};
    for (int k = 1; k < 7; ++k)
      if (cudaLibraryGetKernel(&h->k_lat_ens_a[k], h->lib, names[k]) != cudaSuccess) h->k_lat_ens_a[k] = nullptr;
  }
  if (cudaLibraryGetKernel(&h->k_tiled2d_h[1], h->lib, "odeintr_lattice_rk4_tiled2d_h1") != cudaSuccess) h->k_tiled2d_h[1] = nullptr;
  if (cudaLibraryGetKernel(&h->k_tiled2d_h[2], h->lib, "odeintr_lattice_rk4_tiled2d_h2") != cudaSuccess) h->k_tiled2d_h[2] = nullptr;
  h->k_tiled2d = h->k_tiled2d_h[1];
  if (cudaLibraryGetKernel(&h->k_dopri5_tiled, h->lib, "odeintr_lattice_dopri5_tiled") != cudaSuccess)
    h->k_dopri5_tiled = nullptr;
  if (cudaLibraryGetKernel(&h->k_dopri5_tiled_h[1], h->lib, "odeintr_lattice_dopri5_tiled_h1") != cudaSuccess)
    h->k_dopri5_tiled_h[1] = nullptr;
  if (cudaLibraryGetKernel(&h->k_dopri5_tiled_h[2], h->lib, "odeintr_lattice_dopri5_tiled_h2") != cudaSuccess)
    h->k_dopri5_tiled_h[2] = nullptr;
  if (cudaLibraryGetKernel(&h->k_warp5_h[1], h->lib, "odeintr_lattice_dopri5_warp_h1") != cudaSuccess) h->k_warp5_h[1] = nullptr;
  if (cudaLibraryGetKernel(&h->k_warp5_h[2], h->lib, "odeintr_lattice_dopri5_warp_h2") != cudaSuccess) h->k_warp5_h[2] = nullptr;
  if (cudaLibraryGetKernel(&h->k_warp5_info, h->lib, "odeintr_lattice_dopri5_warp_info") != cudaSuccess) h->k_warp5_info = nullptr;
  if (cudaLibraryGetKernel(&h->k_stream5, h->lib, "odeintr_lattice_dopri5_stream2d") != cudaSuccess) h->k_stream5 = nullptr;
  if (cudaLibraryGetKernel(&h->k_stream5_info, h->lib, "odeintr_lattice_dopri5_stream2d_info") != cudaSuccess) h->k_stream5_info = nullptr;
  if (cudaLibraryGetKernel(&h->k_stream2d, h->lib, "odeintr_lattice_rk4_stream2d") != cudaSuccess) h->k_stream2d = nullptr;
  if (cudaLibraryGetKernel(&h->k_stream2d_info, h->lib, "odeintr_lattice_stream2d_info") != cudaSuccess) h->k_stream2d_info = nullptr;
  if (cudaLibraryGetKernel(&h->k_info, h->lib, "odeintr_lattice_info") != cudaSuccess) h->k_info = nullptr;
  if (cudaLibraryGetKernel(&h->k_serial, h->lib, "odeintr_lattice_serial_rhs") != cudaSuccess) h->k_serial = nullptr;
  if (cudaLibraryGetKernel(&h->k_probe, h->lib, "odeintr_lattice_probe") != cudaSuccess) h->k_probe = nullptr;
  (void)cudaGetLastError();
  if (!h->k_plain && !h->k_adaptive && !h->k_lattice) {
    odeintr_destroy(h);
    return fail(ODEINTR_E_COMPILE, "generated translation unit defines no odeintr kernel");
  }
  auto max_block = [&](cudaKernel_t k, int& block) {
    cudaFuncAttributes fa;
    if (k && cudaFuncGetAttributes(&fa, (const void*)k) == cudaSuccess) block = std::min(block, fa.maxThreadsPerBlock);
  };
  if (const char* b = std::getenv("ODEINTR_BLOCK")) { h->block_plain = std::atoi(b); h->block_adaptive = std::atoi(b); }
  max_block(h->k_plain, h->block_plain);
  max_block(h->k_adaptive, h->block_adaptive);
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  h->own_stream = true;
  if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
  if (h->k_lattice) {
    const int rc = lattice_query_info(h);
    if (rc) {
      odeintr_destroy(h);
      return rc;
    }
  }
  *out = h;
  return ODEINTR_OK;
}

int odeintr_destroy(odeintr_system_t h) {
  if (!h) return ODEINTR_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->x.release(); h->pars.release(); h->out.release(); h->stage.release(); h->t_rec.release();
  h->seg_nfull.release(); h->seg_hrem.release(); h->seg_t.release(); h->seg_obs.release(); h->times_dev.release();
  h->accepted.release(); h->rejected.release(); h->status.release(); h->rows_inst.release(); h->fail_flags.release(); h->rows_min.release();
  h->work_counter.release();
  h->perm.release(); h->perm_iota.release(); h->order_keys.release(); h->order_keys_sorted.release(); h->sort_temp.release();
  h->lat_buf_a.release(); h->lat_buf_b.release(); h->lat_buf_acc.release();
  h->lat_work.release(); h->lat_err.release(); h->lat_flag.release(); h->pipe_bcast.release(); h->lat_negzero.release();
  for (double* v : h->lat_pool) cudaFree(v);
  h->lat_pool.clear();
  for (int s = 0; s < 3; ++s) {
    h->pipe_x[s].release(); h->pipe_out[s].release(); h->pipe_pars[s].release();
    if (h->pipe_done[s]) cudaEventDestroy(h->pipe_done[s]);
    if (h->pipe_stream[s]) cudaStreamDestroy(h->pipe_stream[s]);
  }
  if (h->ev_mid) cudaEventDestroy(h->ev_mid);
  if (h->ev_side) cudaEventDestroy(h->ev_side);
  if (h->ev_edges) cudaEventDestroy(h->ev_edges);
  if (h->ev_xchg) cudaEventDestroy(h->ev_xchg);
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (h->lib) cudaLibraryUnload(h->lib);
  delete h;
  return ODEINTR_OK;
}

const char* odeintr_compile_log(odeintr_system_t h) { return h ? h->log.c_str() : ""; }

int odeintr_set_device(int device) {
  CU_TRY(cudaSetDevice(device));
  return ODEINTR_OK;
}

int odeintr_set_stream(odeintr_system_t h, void* cuda_stream) {
  int rc = activate(h);
  if (rc) return rc;
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  if (cuda_stream) {
    h->stream = (cudaStream_t)cuda_stream;
  } else {
    CU_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  return ODEINTR_OK;
}

int odeintr_set_max_tries(odeintr_system_t h, long long max_tries) {
  if (!h || max_tries < 1) return fail(ODEINTR_E_INVALID, "max_tries must be >= 1");
  h->max_tries = max_tries;
  return ODEINTR_OK;
}

int odeintr_set_instance_order(odeintr_system_t h, int mode, const int* perm, size_t n) {
  int rc = activate(h);
  if (rc) return rc;
  if (mode == ODEINTR_ORDER_INDEX || mode == ODEINTR_ORDER_COST_PROXY) {
    h->order_mode = mode;
    return ODEINTR_OK;
  }
  if (mode != ODEINTR_ORDER_PERMUTATION || !perm || n == 0 || n >= (1ull << 31))
    return fail(ODEINTR_E_INVALID, "Invalid instance order");
  std::vector<unsigned char> seen(n, 0);
  for (size_t i = 0; i < n; ++i) {
    if (perm[i] < 0 || (size_t)perm[i] >= n || seen[(size_t)perm[i]]) return fail(ODEINTR_E_INVALID, "Invalid instance order: not a permutation");
    seen[(size_t)perm[i]] = 1;
  }
  CU_TRY(h->perm.reserve(n));
  if ((rc = upload(h, h->perm.p, perm, n * sizeof(int)))) return rc;
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->perm_m = n;
  h->order_mode = mode;
  return ODEINTR_OK;
}

int odeintr_get_instance_order(odeintr_system_t h, int* perm, size_t n) {
  int rc = activate(h);
  if (rc) return rc;
  if (!perm || h->perm.cap < n || n == 0) return fail(ODEINTR_E_STATE, "no instance order has been computed");
  CU_TRY(cudaMemcpyAsync(perm, h->perm.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

int odeintr_synchronize(odeintr_system_t h) {
  int rc = activate(h);
  if (rc) return rc;
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

// ---- state and parameters --------------------------------------------------------------------------
int odeintr_set_state(odeintr_system_t h, const double* x, size_t n_values, size_t n_inst) {
  int rc = activate(h);
  if (rc) return rc;
  if (!x || n_inst == 0 || n_values != (size_t)h->N * n_inst) return fail(ODEINTR_E_INVALID, "Invalid initial state");
  CU_TRY(h->x.reserve(n_values));
  h->m = n_inst;
  if ((rc = upload(h, h->x.p, x, n_values * sizeof(double)))) return rc;
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

int odeintr_get_state(odeintr_system_t h, double* x, size_t n_values) {
  int rc = activate(h);
  if (rc) return rc;
  if (!x || h->m == 0 || n_values != (size_t)h->N * h->m) return fail(ODEINTR_E_INVALID, "Invalid state buffer");
  CU_TRY(cudaMemcpyAsync(x, h->x.p, n_values * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

int odeintr_set_params(odeintr_system_t h, const double* p, size_t n_values, size_t n_inst) {
  int rc = activate(h);
  if (rc) return rc;
  if (h->P == 0) return fail(ODEINTR_E_INVALID, "system has no run-time parameters");
  if (!p || n_inst == 0 || n_values != (size_t)h->P * n_inst) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  CU_TRY(h->pars.reserve(n_values));
  if ((rc = upload(h, h->pars.p, p, n_values * sizeof(double)))) return rc;
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->par_sets = n_inst;
  return ODEINTR_OK;
}

int odeintr_get_params(odeintr_system_t h, double* p, size_t n_values) {
  int rc = activate(h);
  if (rc) return rc;
  if (!p || h->par_sets == 0 || n_values != (size_t)h->P * h->par_sets)
    return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  CU_TRY(cudaMemcpyAsync(p, h->pars.p, n_values * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

size_t odeintr_n_instances(odeintr_system_t h) { return h ? h->m : 0; }
size_t odeintr_n_param_sets(odeintr_system_t h) { return h ? h->par_sets : 0; }

int odeintr_fill_state_uniform(odeintr_system_t h, size_t n_inst, unsigned long long seed,
                               unsigned long long first_instance, const double* lo, const double* hi) {
  int rc = activate(h);
  if (rc) return rc;
  if (!lo || !hi || n_inst == 0) return fail(ODEINTR_E_INVALID, "Invalid initial state");
  CU_TRY(h->x.reserve((size_t)h->N * n_inst));
  CU_TRY(h->stage.reserve(2 * (size_t)std::max(h->N, h->P)));
  std::vector<double> lohi(2 * h->N);
  for (int v = 0; v < h->N; ++v) { lohi[v] = lo[v]; lohi[h->N + v] = hi[v]; }
  if ((rc = upload(h, h->stage.p, lohi.data(), lohi.size() * sizeof(double)))) return rc;
  odeintr_b200::launch_fill_state_uniform(h->x.p, (long long)n_inst, (long long)n_inst, h->N, seed, first_instance,
                                          h->stage.p, h->stream);
  g_total_launches += 1;
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->m = n_inst;
  return ODEINTR_OK;
}

int odeintr_fill_params_linspace(odeintr_system_t h, size_t n_inst, unsigned long long first_instance,
                                 unsigned long long total_instances, const double* lo, const double* hi) {
  int rc = activate(h);
  if (rc) return rc;
  if (h->P == 0) return fail(ODEINTR_E_INVALID, "system has no run-time parameters");
  if (!lo || !hi || n_inst == 0) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  CU_TRY(h->pars.reserve((size_t)h->P * n_inst));
  CU_TRY(h->stage.reserve(2 * (size_t)std::max(h->N, h->P)));
  std::vector<double> lohi(2 * h->P);
  for (int k = 0; k < h->P; ++k) { lohi[k] = lo[k]; lohi[h->P + k] = hi[k]; }
  if ((rc = upload(h, h->stage.p, lohi.data(), lohi.size() * sizeof(double)))) return rc;
  odeintr_b200::launch_fill_params_linspace(h->pars.p, (long long)n_inst, (long long)n_inst, h->P, first_instance,
                                            total_instances, h->stage.p, h->stream);
  g_total_launches += 1;
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->par_sets = n_inst;
  return ODEINTR_OK;
}

// ---- integration -----------------------------------------------------------------------------------
int odeintr_integrate_const(odeintr_system_t h, double t0, double t1, double dt, long long obs_stride, int record) {
  int rc = activate(h);
  if (rc) return rc;
  if (!(dt != 0.0) || !std::isfinite(dt) || !std::isfinite(t0) || !std::isfinite(t1))
    return fail(ODEINTR_E_INVALID, "Invalid time specification");
  if (obs_stride < 1) return fail(ODEINTR_E_INVALID, "obs_stride must be >= 1");
  if (h->k_lattice && h->lattice_method) return run_lattice_adaptive(h, ODEINTR_MODE_CONST, t0, t1, dt, nullptr, 0, record != 0);
  if (h->k_plain || h->k_lattice) {
    plain_schedule sc;
    schedule_const(sc, t0, t1, dt, obs_stride, record != 0);
    return h->k_lattice ? run_lattice(h, sc) : run_plain(h, sc);
  }
  if (h->k_adaptive) {
    if (obs_stride != 1) return fail(ODEINTR_E_INVALID, "obs_stride applies to fixed-step steppers only");
    return run_adaptive(h, ODEINTR_MODE_CONST, t0, t1, dt, nullptr, 0, record != 0, false);
  }
  return fail(ODEINTR_E_STATE, "integrate_const: no kernel for this stepper family");
}

int odeintr_integrate_times(odeintr_system_t h, const double* times, size_t n_times, double dt) {
  int rc = activate(h);
  if (rc) return rc;
  if (!times || n_times == 0) return fail(ODEINTR_E_INVALID, "Invalid time specification");
  if (!(dt != 0.0) || !std::isfinite(dt)) return fail(ODEINTR_E_INVALID, "Invalid time specification");
  if (h->k_lattice && h->lattice_method)
    return run_lattice_adaptive(h, ODEINTR_MODE_TIMES, times[0], times[n_times - 1], dt, times, n_times, true);
  if (h->k_plain || h->k_lattice) {
    plain_schedule sc;
    schedule_times(sc, times, n_times, dt);
    return h->k_lattice ? run_lattice(h, sc) : run_plain(h, sc);
  }
  if (h->k_adaptive) return run_adaptive(h, ODEINTR_MODE_TIMES, times[0], times[n_times - 1], dt, times, n_times, true, false);
  return fail(ODEINTR_E_STATE, "integrate_times: no kernel for this stepper family");
}

int odeintr_integrate_adaptive(odeintr_system_t h, double t0, double t1, double dt, int record) {
  int rc = activate(h);
  if (rc) return rc;
  if (!(dt != 0.0) || !std::isfinite(dt) || !std::isfinite(t0) || !std::isfinite(t1))
    return fail(ODEINTR_E_INVALID, "Invalid time specification");
  if (h->k_lattice && h->lattice_method) return run_lattice_adaptive(h, ODEINTR_MODE_ADAPTIVE, t0, t1, dt, nullptr, 0, record != 0);
  if (h->k_plain || h->k_lattice) {
    plain_schedule sc;
    schedule_adaptive(sc, t0, t1, dt, record != 0);
    return h->k_lattice ? run_lattice(h, sc) : run_plain(h, sc);
  }
  if (h->k_adaptive) return run_adaptive(h, ODEINTR_MODE_ADAPTIVE, t0, t1, dt, nullptr, 0, record != 0, false);
  return fail(ODEINTR_E_STATE, "integrate_adaptive: no kernel for this stepper family");
}

int odeintr_integrate_const_host(odeintr_system_t h, const double* x, size_t n_inst, const double* pars,
                                 size_t par_sets, double t0, double t1, double dt, long long obs_stride,
                                 double* out, size_t out_values, double* x_final, size_t chunk) {
  int rc = activate(h);
  if (rc) return rc;
  if (!(dt != 0.0) || !std::isfinite(dt) || !std::isfinite(t0) || !std::isfinite(t1))
    return fail(ODEINTR_E_INVALID, "Invalid time specification");
  if (obs_stride < 1) return fail(ODEINTR_E_INVALID, "obs_stride must be >= 1");
  if (!h->k_plain) return fail(ODEINTR_E_STATE, "integrate_const_host: no kernel for this stepper family");
  plain_schedule sc;
  schedule_const(sc, t0, t1, dt, obs_stride, out != nullptr);
  if (out && out_values != sc.rec_t.size() * (size_t)h->N * n_inst) return fail(ODEINTR_E_INVALID, "Invalid output buffer");
  return run_plain_host(h, sc, x, n_inst, pars, par_sets, out, x_final, chunk);
}

size_t odeintr_const_rows(double t0, double t1, double dt, long long obs_stride) {
  if (!(dt != 0.0) || obs_stride < 1) return 0;
  const long long n = const_step_count(t0, t1, dt);
  return (size_t)((n + obs_stride - 1) / obs_stride) + 1;
}

// ---- slab-sharded large systems (one process per GPU) -------------------------------------------------------
int odeintr_lattice_shard(odeintr_system_t h, long long cells_total, long long first_cell, long long n_local,
                          int fields, int halo, double* x_dev) {
  int rc = activate(h);
  if (rc) return rc;
  if (!h->k_lattice) return fail(ODEINTR_E_STATE, "system was not compiled with compile_sys_openmp");
  if (cells_total < 1 || n_local < 1 || first_cell < 0 || first_cell + n_local > cells_total || fields < 1 || halo < 0)
    return fail(ODEINTR_E_INVALID, "Invalid shard specification");
  if ((long long)fields * cells_total != (long long)h->N) return fail(ODEINTR_E_INVALID, "fields * cells_total must equal sys_dim");
  if (!x_dev) return fail(ODEINTR_E_INVALID, "null state pointer");
  h->sh_tiled = !h->lattice_method && slab_tiled_possible(h, halo) && cells_total == h->lat_cells_ct && tiled_configure(h, halo, fields);
  h->sh_steps_per_exchange = h->sh_tiled ? h->sh_interval : 1;
  const long long ghost = odeintr_lattice_ghost_width(h, halo);
  if (n_local + 2 * ghost > cells_total && n_local != cells_total)
    return fail(ODEINTR_E_INVALID, "slab too small for the halo width");
  if (n_local < ghost) return fail(ODEINTR_E_INVALID, "slab too small for the halo width");
  h->sh_total = cells_total; h->sh_c0 = first_cell; h->sh_n = n_local; h->sh_fields = fields; h->sh_halo = halo;
  h->sh_ghost = ghost;
  h->sh_pad = odeintr_lattice_pad_width(h, halo);
  h->sh_pitch = odeintr_lattice_pitch(h, halo, n_local);
  h->sh_x = x_dev;
  const size_t total = (size_t)fields * (size_t)h->sh_pitch;
  if (!h->lattice_method) {  // (adaptive steppers keep their work vectors in lat_work)
    CU_TRY(h->lat_buf_a.reserve(total));
    if (!h->sh_tiled) { CU_TRY(h->lat_buf_b.reserve(total)); CU_TRY(h->lat_buf_acc.reserve(total)); }
    else if (h->sh_steps_per_exchange >= 2 && h->k_warp && h->warp_blocks > 0) CU_TRY(h->lat_buf_acc.reserve(total));  // fused step pairs
  }
  CU_TRY(h->lat_flag.reserve(8));
  CU_TRY(cudaMemsetAsync(h->lat_flag.p, 0, sizeof(int), h->stream));
  h->sh_primed = false;
  h->sharded = true;
  return ODEINTR_OK;
}

long long odeintr_lattice_ghost_width(odeintr_system_t h, int halo) {
  if (!h) return 0;
  // adaptive dopri5 (rk5_i / rk5_a): an attempt forms six neighbour-dependent stages; exchanged after every accepted step
  if (h->lattice_method) return 7LL * halo;
  return (long long)halo * h->lattice_stages * (slab_tiled_possible(h, halo) ? h->sh_interval : 1);
}

int odeintr_lattice_set_exchange_interval(odeintr_system_t h, int steps) {
  if (!h) return fail(ODEINTR_E_INVALID, "null system handle");
  if (steps < 1 || steps > 64) return fail(ODEINTR_E_INVALID, "steps per halo exchange must be in [1, 64]");
  if (h->sharded) return fail(ODEINTR_E_STATE, "set the exchange interval before odeintr_lattice_shard");
  h->sh_interval = steps;
  return ODEINTR_OK;
}

// cells reserved in front of (and behind) the owned cells of every field: the ghost width rounded up to a
// multiple of 16 doubles so that local cell 0 of every field is 128-byte aligned
long long odeintr_lattice_pad_width(odeintr_system_t h, int halo) {
  const long long g = odeintr_lattice_ghost_width(h, halo);
  return (g + 15) / 16 * 16;
}

// doubles per field in the caller's state buffer: pad + n_local + pad, rounded up to a multiple of 16
long long odeintr_lattice_pitch(odeintr_system_t h, int halo, long long n_local) {
  const long long pad = odeintr_lattice_pad_width(h, halo);
  return (pad + n_local + pad + 15) / 16 * 16;
}

namespace {
// one rk4 step of the slab through the tiled kernels: `cur` -> `other` (both laid out like the caller's buffer), the
// ghost cells of `cur` being valid out to ghost - phase * 4 * halo cells (phase = steps since the last exchange).
// The cells of a step split into the MIDDLE [c0 + g, c0 + n - g), which in the first step after an exchange depends on
// owned cells only, and the EDGES on either side of it, which need the ghost cells and, in the last step of a round,
// are what the next exchange sends: odeintr_lattice_shard_run orders the two parts around the exchange so that the
// ring transfer travels while the middle is being computed.
enum { SLAB_ALL = 0, SLAB_EDGES = 1, SLAB_MIDDLE = 2 };
int slab_tiled_step(odeintr_system_t h, double* cur, double* other, double t, double dt, int phase, int part) {
  const long long g = h->sh_ghost, pad = h->sh_pad, margin = 4 * (long long)h->sh_halo;
  h->lat_x = cur + pad; h->lat_a = other + pad;
  lattice_window w = {h->sh_n, h->sh_total, h->sh_c0, g, h->sh_pitch, 1, 0, 0};
  w.lo = h->sh_c0 - g + margin * (phase + 1);
  w.hi = h->sh_c0 + h->sh_n + g - margin * (phase + 1);
  const long long e_lo = h->sh_c0 + g, e_hi = h->sh_c0 + h->sh_n - g;
  if (part == SLAB_ALL || e_hi - e_lo < 1024) {  // (slabs too small to split: everything counts as edge)
    if (part == SLAB_MIDDLE) return ODEINTR_OK;
    return lattice_step_halo(h, w, t, dt, h->sh_halo);
  }
  odeintr_lattice_tiled_args ta;
  tiled_args_init(h, w, h->lat_x, h->lat_a, t, dt, ta);
  const long long glo = h->lat_start > 0 ? h->lat_start : 0;
  const long long ghi = h->lat_bound < ta.n_total ? h->lat_bound : ta.n_total;
  if (part == SLAB_EDGES) return tiled_general(h, ta, w.lo, e_lo, e_hi, w.hi);
  return tiled_produce(h, ta, e_lo, e_hi, glo, ghi, true);
}

// Two consecutive rk4 steps of the slab (phases `phase` and `phase + 1` of one round) as ONE pass of the warp kernel over
// the middle of the slab (x -> x'' in registers, 8 B per cell and step: what lattice_step_pair does on one GPU); the
// edges -- everything between the window of the step and the middle -- take the general kernel twice through `mid`.
// part = SLAB_EDGES: both edge steps (the second one's results are what an exchange after the pair sends);
// part = SLAB_MIDDLE: the fused pass.  slab_pair_possible decides once per run.
bool slab_pair_possible(odeintr_system_t h, long long* ilo_out, long long* ihi_out, bool* tma_out, const double* cur,
                        const double* other) {
  const char* knob = std::getenv("ODEINTR_FUSE");
  if (knob && std::atoi(knob) == 0) return false;
  if (h->lattice_stages != 4 || !h->sh_tiled || !h->k_tiled || !h->k_warp || h->warp_blocks <= 0 || h->k_serial ||
      h->sh_steps_per_exchange < 2 || std::getenv("ODEINTR_NO_WARP") || std::getenv("ODEINTR_TILED_GENERAL_ONLY"))
    return false;
  if (h->lat_cells_ct != h->sh_total) return false;
  const long long g = h->sh_ghost, m1 = 4 * (long long)h->sh_halo;
  const long long glo = h->lat_start > 0 ? h->lat_start : 0;
  const long long ghi = h->lat_bound < h->sh_total ? h->lat_bound : h->sh_total;
  const long long e_lo = h->sh_c0 + g, e_hi = h->sh_c0 + h->sh_n - g;
  // the fused pass reads 2 * m1 (+ one radius: it evaluates its outermost cells too) beyond what it produces: owned
  // cells only, so that it can start before the ghost cells arrive, and inside the user loop's range
  long long ilo = std::max(std::max(e_lo, h->sh_c0 + 2 * m1 + h->sh_halo), glo + 2 * m1 + h->sh_halo);
  long long ihi = std::min(std::min(e_hi, h->sh_c0 + h->sh_n - 2 * m1 - h->sh_halo), ghi - 2 * m1 - h->sh_halo);
  const bool tma = reinterpret_cast<uintptr_t>(cur) % 16 == 0 && reinterpret_cast<uintptr_t>(other) % 16 == 0 &&
                   (h->sh_fields == 1 || h->sh_pitch % 2 == 0) && !std::getenv("ODEINTR_NO_TMA");
  if (tma) {
    if ((ilo - h->sh_c0) & 1) ++ilo;
    if ((ihi - h->sh_c0) & 1) --ihi;
  }
  const long long P = 32LL * h->warp_cpt - 16LL * h->warp_halo;
  if (P < 32 || ihi - ilo < 4 * P) return false;
  *ilo_out = ilo; *ihi_out = ihi; *tma_out = tma;
  return true;
}

int slab_tiled_pair(odeintr_system_t h, double* cur, double* mid, double* other, double t, double t2, double dt, int phase,
                    int part, long long ilo, long long ihi, bool tma) {
  const long long g = h->sh_ghost, pad = h->sh_pad, m1 = 4 * (long long)h->sh_halo;
  lattice_window w = {h->sh_n, h->sh_total, h->sh_c0, g, h->sh_pitch, 1, 0, 0};
  const long long lo1 = h->sh_c0 - g + m1 * (phase + 1), hi1 = h->sh_c0 + h->sh_n + g - m1 * (phase + 1);
  const long long lo2 = lo1 + m1, hi2 = hi1 - m1;
  int rc;
  if (part == SLAB_EDGES) {
    // first step of the edges: cur -> mid on everything the second step of the edges reads
    odeintr_lattice_tiled_args e1;
    tiled_args_init(h, w, cur + pad, mid + pad, t, dt, e1);
    if ((rc = tiled_general(h, e1, lo1, std::min(ilo + m1, hi1), std::max(ihi - m1, lo1), hi1))) return rc;
    odeintr_lattice_tiled_args e2;
    tiled_args_init(h, w, mid + pad, other + pad, t2, dt, e2);
    return tiled_general(h, e2, lo2, ilo, ihi, hi2);
  }
  odeintr_lattice_tiled_args ta;
  tiled_args_init(h, w, cur + pad, other + pad, t, dt, ta);
  ta.lo = ilo; ta.hi = ihi; ta.sharded = 0; ta.tma = tma ? 1 : 0; ta.nsub = 2; ta.t2 = t2;
  const long long P = 32LL * h->warp_cpt - 16LL * h->warp_halo;
  const long long warp_tiles = (ihi - ilo + P - 1) / P;
  const unsigned grid = (unsigned)std::min<long long>((warp_tiles + 7) / 8, (long long)h->sm_count * h->warp_blocks);
  return tiled_launch(h, h->k_warp, grid, h->warp_smem, ta);
}

int slab_check_reach(odeintr_system_t h) {
  int flag = 0;
  CU_TRY(cudaMemcpyAsync(&flag, h->lat_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (flag)
    return fail(ODEINTR_E_INVALID, "the system definition reads x further than the declared halo (" +
                                       std::to_string(h->sh_halo) + " cells) from the cell it computes");
  return ODEINTR_OK;
}

int slab_prime(odeintr_system_t h) {
  // work buffers start as copies of the state so that entries no loop writes keep their value
  if (h->sh_primed) return ODEINTR_OK;
  const size_t total = (size_t)h->sh_fields * (size_t)h->sh_pitch;
  CU_TRY(cudaMemcpyAsync(h->lat_buf_a.p, h->sh_x, total * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (h->sh_tiled && h->lat_buf_acc.p && h->lat_buf_acc.cap >= total)  // intermediate state of fused step pairs
    CU_TRY(cudaMemcpyAsync(h->lat_buf_acc.p, h->sh_x, total * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (!h->sh_tiled) {
    CU_TRY(cudaMemcpyAsync(h->lat_buf_b.p, h->sh_x, total * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->lat_buf_acc.p, h->sh_x, total * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  h->sh_primed = true;
  return ODEINTR_OK;
}
}  // namespace

int odeintr_lattice_shard_step(odeintr_system_t h, double t, double dt) {
  int rc = activate(h);
  if (rc) return rc;
  if (!h->sharded) return fail(ODEINTR_E_STATE, "odeintr_lattice_shard has not been called");
  if (h->P && h->par_sets != 1) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  if (h->lattice_stages != 4) return fail(ODEINTR_E_STATE, "sharded stepping is provided for rk4");
  if ((rc = slab_prime(h))) return rc;
  h->last_launches = 0;
  if (h->sh_tiled) {
    // single-step entry point: the caller has just exchanged the halos and finds the new state in its own buffer
    const size_t total = (size_t)h->sh_fields * (size_t)h->sh_pitch;
    if ((rc = slab_tiled_step(h, h->sh_x, h->lat_buf_a.p, t, dt, 0, SLAB_ALL))) return rc;
    CU_TRY(cudaMemcpyAsync(h->sh_x, h->lat_buf_a.p, total * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return slab_check_reach(h);
  }
  const long long g = h->sh_ghost, pad = h->sh_pad;
  h->lat_x = h->sh_x + pad; h->lat_a = h->lat_buf_a.p + pad; h->lat_b = h->lat_buf_b.p + pad;
  h->lat_acc = h->lat_buf_acc.p + pad;
  lattice_window w = {h->sh_n, h->sh_total, h->sh_c0, g, h->sh_pitch, 1, 0, 0};
  if ((rc = lattice_step_halo(h, w, t, dt, h->sh_halo))) return rc;
  return ODEINTR_OK;
}

int odeintr_lattice_set_halo(odeintr_system_t h, long long halo, int fields) {
  int rc = activate(h);
  if (rc) return rc;
  if (!h->k_lattice) return fail(ODEINTR_E_STATE, "system was not compiled with compile_sys_openmp");
  h->tiled_on = h->tiled2d_on = h->tiled_auto = h->tiled_pending = false;
  if (halo < 0) {  // measured at the first integrate call, when the parameters are in place
    h->tiled_pending = true;
    return ODEINTR_OK;
  }
  if (fields < 1 || (long long)h->N % fields != 0) return fail(ODEINTR_E_INVALID, "Invalid halo specification");
  if (h->lat_cells_ct > 0 && (long long)h->N / fields != h->lat_cells_ct)
    return fail(ODEINTR_E_INVALID, "Invalid halo specification: the system writes " +
                                       std::to_string((long long)h->N / h->lat_cells_ct) + " field(s)");
  if (halo == 0) return ODEINTR_OK;  // staged kernels
  if (h->lat_side > 0) {  // M x M lattice
    if (h->lattice_stages != 4) return fail(ODEINTR_E_STATE, "the tiled kernel is an rk4 step");
    if (!tiled2d_configure(h, halo)) return fail(ODEINTR_E_INVALID, "halo too wide for the shared-memory tile");
    h->tiled2d_on = true;
    return ODEINTR_OK;
  }
  if (!h->k_tiled) return fail(ODEINTR_E_STATE, "the generated translation unit has no tiled kernel");
  if (h->lattice_stages != 4) return fail(ODEINTR_E_STATE, "the tiled kernel is an rk4 step");
  if ((long long)h->N / fields >= (1LL << 31))
    return fail(ODEINTR_E_INVALID, "the tiled kernel indexes cells with 32 bits: at most 2^31 - 1 cells per field");
  if (!tiled_configure(h, halo, fields)) return fail(ODEINTR_E_INVALID, "halo too wide for the shared-memory tile");
  h->lattice_cells = (long long)h->N / fields;
  h->tiled_on = true;
  return ODEINTR_OK;
}

long long odeintr_lattice_get_halo(odeintr_system_t h) {
  if (!h || activate(h)) return -1;
  if (h->tiled_pending && tiled_autodetect(h)) return -1;
  return (h->tiled_on || h->tiled2d_on) ? h->lattice_halo : -1;
}

int odeintr_nccl_unique_id(void* id_out, size_t cap) {
  if (!id_out || cap < sizeof(ncclUniqueId)) return fail(ODEINTR_E_INVALID, "unique-id buffer too small (128 bytes)");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return ODEINTR_OK;
}

int odeintr_lattice_comm_init(odeintr_system_t h, int world, int rank, const void* id_bytes, size_t id_size) {
  int rc = activate(h);
  if (rc) return rc;
  if (world < 1 || rank < 0 || rank >= world) return fail(ODEINTR_E_INVALID, "invalid rank / world size");
  h->comm_world = world; h->comm_rank = rank;
  if (world == 1) return ODEINTR_OK;
  if (!id_bytes || id_size < sizeof(ncclUniqueId)) return fail(ODEINTR_E_INVALID, "missing NCCL unique id");
  if ((rc = load_nccl())) return rc;
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, sizeof(id));
  NCCL_TRY(g_nccl.CommInitRank(&h->comm, world, id, rank));
  return ODEINTR_OK;
}


int odeintr_lattice_shard_run(odeintr_system_t h, long long first_step, long long n_steps, double t0, double dt) {
  int rc = activate(h);
  if (rc) return rc;
  if (!h->sharded) return fail(ODEINTR_E_STATE, "odeintr_lattice_shard has not been called");
  if (n_steps < 0) return fail(ODEINTR_E_INVALID, "negative step count");
  if (h->lattice_stages != 4) return fail(ODEINTR_E_STATE, "sharded stepping is provided for rk4");
  if (h->P && h->par_sets != 1) return fail(ODEINTR_E_INVALID, "Invalid parameter vector");
  if ((rc = slab_prime(h))) return rc;
  if ((rc = begin_timing(h))) return rc;
  if (h->sh_tiled) {
    // the state ping-pongs between the caller's buffer and ours; ghost cells are wide enough for
    // sh_steps_per_exchange steps (each consumes 4 * halo of them), so the ring exchange runs once per round.
    // The exchange travels on a second stream: the last step of a round computes its edges first (the cells the
    // exchange sends), hands them over and goes on with the middle of the slab; the first step of the next round
    // starts with its middle and waits for the ghost cells only before its edges.
    const bool overlap = !std::getenv("ODEINTR_NO_OVERLAP");
    if (overlap && !h->stream2) {
      int lo_prio = 0, hi_prio = 0;
      CU_TRY(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
      CU_TRY(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hi_prio));
      CU_TRY(cudaEventCreateWithFlags(&h->ev_edges, cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&h->ev_xchg, cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&h->ev_mid, cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
    }
    double* cur = h->sh_x;
    double* other = h->lat_buf_a.p;
    const int spe = h->sh_steps_per_exchange;
    long long launches = 0;
    auto exchange_async = [&](double* buf) -> int {  // after everything queued on the main stream so far
      CU_TRY(cudaEventRecord(h->ev_edges, h->stream));
      CU_TRY(cudaStreamWaitEvent(h->stream2, h->ev_edges, 0));
      int erc = lattice_exchange(h, buf, h->stream2);
      if (erc) return erc;
      CU_TRY(cudaEventRecord(h->ev_xchg, h->stream2));
      return ODEINTR_OK;
    };
    long long p_ilo = 0, p_ihi = 0;
    bool p_tma = false;
    bool have_side = false, have_mid = false;  // the second stream carries edges the main stream has not waited for / vice versa
    double* mid = h->lat_buf_acc.p;
    const bool pairs = mid && h->lat_buf_acc.cap >= (size_t)h->sh_fields * (size_t)h->sh_pitch &&
                       slab_pair_possible(h, &p_ilo, &p_ihi, &p_tma, cur, other);
    for (long long s = 0; s < n_steps; ++s) {
      const int phase = (int)(s % spe);
      const double t = t0 + static_cast<double>(first_step + s) * dt;
      h->last_launches = 0;
      if (pairs && phase + 1 < spe && s + 1 < n_steps) {
        // two steps of one round in one pass over the middle of the slab (the state between them stays in registers)
        const double t2 = t0 + static_cast<double>(first_step + s + 1) * dt;
        const bool first_of_round = phase == 0;
        const bool feeds_exchange = phase + 1 == spe - 1 && s + 2 < n_steps;  // the step after the pair starts a round
        if (!overlap) {
          if (first_of_round && (rc = lattice_exchange(h, cur, h->stream))) return rc;
          if ((rc = slab_tiled_pair(h, cur, mid, other, t, t2, dt, phase, SLAB_EDGES, p_ilo, p_ihi, p_tma))) return rc;
          if ((rc = slab_tiled_pair(h, cur, mid, other, t, t2, dt, phase, SLAB_MIDDLE, p_ilo, p_ihi, p_tma))) return rc;
        } else {
          // The edges of a pair (two launches of a few blocks, then -- at the end of a round -- the ring exchange) travel
          // on the second stream BESIDE the fused pass over the middle on the main stream: each side waits only for the
          // other side's previous pair (the middle reads a margin of what the edges produced and vice versa).  At 2^25
          // cells per GPU the two edge launches were 10 % of a step when they ran in line with the middle.
          if (s == 0 && (rc = exchange_async(cur))) return rc;  // (the second stream now follows the main one)
          (void)first_of_round;  // the exchange that feeds this round was queued on the second stream ahead of these edges
          if (have_side) CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_side, 0));
          if (have_mid) CU_TRY(cudaStreamWaitEvent(h->stream2, h->ev_mid, 0));
          {
            cudaStream_t main_stream = h->stream;
            h->stream = h->stream2;
            rc = slab_tiled_pair(h, cur, mid, other, t, t2, dt, phase, SLAB_EDGES, p_ilo, p_ihi, p_tma);
            h->stream = main_stream;
            if (rc) return rc;
          }
          CU_TRY(cudaEventRecord(h->ev_side, h->stream2));  // (before the exchange: the next middle does not wait for it)
          have_side = true;
          if (feeds_exchange) {
            if ((rc = lattice_exchange(h, other, h->stream2))) return rc;
            CU_TRY(cudaEventRecord(h->ev_xchg, h->stream2));
          }
          if ((rc = slab_tiled_pair(h, cur, mid, other, t, t2, dt, phase, SLAB_MIDDLE, p_ilo, p_ihi, p_tma))) return rc;
          CU_TRY(cudaEventRecord(h->ev_mid, h->stream));
          have_mid = true;
        }
        launches += h->last_launches;
        std::swap(cur, other);
        ++s;
        continue;
      }
      if (have_side) {  // a single step after fused pairs: everything the second stream still carries comes first
        CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_side, 0));
        have_side = false;
      }
      if (!overlap) {
        if (phase == 0 && (rc = lattice_exchange(h, cur, h->stream))) return rc;
        if ((rc = slab_tiled_step(h, cur, other, t, dt, phase, SLAB_ALL))) return rc;
      } else {
        const bool first_of_round = phase == 0;
        const bool feeds_exchange = phase == spe - 1 && s + 1 < n_steps;  // the next step starts a round
        if (s == 0 && (rc = exchange_async(cur))) return rc;
        if (first_of_round) {
          if ((rc = slab_tiled_step(h, cur, other, t, dt, phase, SLAB_MIDDLE))) return rc;
          CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_xchg, 0));
          if ((rc = slab_tiled_step(h, cur, other, t, dt, phase, SLAB_EDGES))) return rc;
          if (feeds_exchange && (rc = exchange_async(other))) return rc;  // one step per round
        } else if (feeds_exchange) {
          if ((rc = slab_tiled_step(h, cur, other, t, dt, phase, SLAB_EDGES))) return rc;
          if ((rc = exchange_async(other))) return rc;
          if ((rc = slab_tiled_step(h, cur, other, t, dt, phase, SLAB_MIDDLE))) return rc;
        } else {
          if ((rc = slab_tiled_step(h, cur, other, t, dt, phase, SLAB_EDGES))) return rc;
          if ((rc = slab_tiled_step(h, cur, other, t, dt, phase, SLAB_MIDDLE))) return rc;
        }
      }
      if (overlap && pairs) {  // pairs that follow let their edges wait for this step
        CU_TRY(cudaEventRecord(h->ev_mid, h->stream));
        have_mid = true;
      }
      launches += h->last_launches;
      std::swap(cur, other);
    }
    if (overlap && h->stream2 && h->ev_side) {  // whatever the second stream still carries belongs to this call
      CU_TRY(cudaEventRecord(h->ev_side, h->stream2));
      CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_side, 0));
    }
    h->last_launches = launches;
    if (cur != h->sh_x) {
      const size_t total = (size_t)h->sh_fields * (size_t)h->sh_pitch;
      CU_TRY(cudaMemcpyAsync(h->sh_x, cur, total * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    if ((rc = end_timing(h))) return rc;
    h->last_steps = n_steps;
    return slab_check_reach(h);
  }
  long long launches = 0;
  for (long long s = 0; s < n_steps; ++s) {
    if ((rc = lattice_exchange(h, h->sh_x, h->stream))) return rc;
    if ((rc = odeintr_lattice_shard_step(h, t0 + static_cast<double>(first_step + s) * dt, dt))) return rc;  // integrate_const time rule
    launches += h->last_launches;
  }
  h->last_launches = launches;
  if ((rc = end_timing(h))) return rc;
  h->last_steps = n_steps;
  return ODEINTR_OK;
}

int odeintr_lattice_shard_get_output(odeintr_system_t h, double* out, size_t n_values) {
  int rc = activate(h);
  if (rc) return rc;
  if (!h->sharded) return fail(ODEINTR_E_STATE, "odeintr_lattice_shard has not been called");
  const size_t rows = h->rows, nl = (size_t)h->sh_n, pitch = (size_t)h->sh_pitch, fields = (size_t)h->sh_fields;
  if (n_values != rows * fields * nl || (!out && n_values)) return fail(ODEINTR_E_INVALID, "Invalid output buffer");
  if (!n_values) return ODEINTR_OK;
  // record rows are whole slabs [field][pitch]; the caller gets the owned cells [row][field][n_local]
  CU_TRY(cudaMemcpy2DAsync(out, nl * sizeof(double), h->out.p + h->sh_pad, pitch * sizeof(double), nl * sizeof(double),
                           rows * fields, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

int odeintr_lattice_exchange(odeintr_system_t h) {
  int rc = activate(h);
  if (rc) return rc;
  if (!h->sharded) return fail(ODEINTR_E_STATE, "odeintr_lattice_shard has not been called");
  return lattice_exchange(h, h->sh_x, h->stream);
}

// ---- observer record -------------------------------------------------------------------------------
int odeintr_reset_observer(odeintr_system_t h) {
  if (!h) return fail(ODEINTR_E_INVALID, "null system handle");
  h->rec_t.clear();
  h->rows = 0;
  h->ragged = false;
  return ODEINTR_OK;
}

int odeintr_release_record(odeintr_system_t h) {
  int rc = activate(h);
  if (rc) return rc;
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->out.release(); h->stage.release(); h->t_rec.release();
  h->rec_t.clear();
  h->rows = 0;
  h->ragged = false;
  return ODEINTR_OK;
}

size_t odeintr_output_rows(odeintr_system_t h) { return h ? h->rows : 0; }

int odeintr_get_output_times(odeintr_system_t h, double* t, size_t n_values) {
  if (!h) return fail(ODEINTR_E_INVALID, "null system handle");
  if (h->ragged) return fail(ODEINTR_E_STATE, "record is ragged; use odeintr_get_output_times_per_instance");
  if (n_values != h->rec_t.size() || (!t && n_values)) return fail(ODEINTR_E_INVALID, "Invalid output buffer");
  if (n_values) std::memcpy(t, h->rec_t.data(), n_values * sizeof(double));
  return ODEINTR_OK;
}

int odeintr_get_output(odeintr_system_t h, double* x, size_t n_values, size_t inst_begin, size_t inst_count) {
  int rc = activate(h);
  if (rc) return rc;
  if (inst_begin + inst_count > h->out_m) return fail(ODEINTR_E_INVALID, "instance range outside the record");
  const size_t rows = h->ragged ? (size_t)h->max_rows : h->rows;
  const size_t need = rows * (size_t)h->N * inst_count;
  if (n_values != need || (!x && need)) return fail(ODEINTR_E_INVALID, "Invalid output buffer");
  if (need == 0) return ODEINTR_OK;
  CU_TRY(h->stage.reserve(need));
  odeintr_b200::launch_gather_output(h->out.p, h->stage.p, (long long)rows, h->N, (long long)h->out_m,
                                     (long long)inst_begin, (long long)inst_count, h->stream);
  g_total_launches += 1;
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpyAsync(x, h->stage.p, need * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

int odeintr_get_output_soa(odeintr_system_t h, double* out, size_t n_values) {
  int rc = activate(h);
  if (rc) return rc;
  const size_t rows = h->ragged ? (size_t)h->max_rows : h->rows;
  const size_t need = rows * (size_t)h->N * h->out_m;
  if (n_values != need || (!out && need)) return fail(ODEINTR_E_INVALID, "Invalid output buffer");
  if (need == 0) return ODEINTR_OK;
  CU_TRY(cudaMemcpyAsync(out, h->out.p, need * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

int odeintr_output_is_ragged(odeintr_system_t h) { return (h && h->ragged) ? 1 : 0; }

int odeintr_get_output_rows_per_instance(odeintr_system_t h, int* rows, size_t n_values) {
  int rc = activate(h);
  if (rc) return rc;
  if (!rows || n_values != h->out_m) return fail(ODEINTR_E_INVALID, "Invalid output buffer");
  if (!h->ragged) {
    for (size_t i = 0; i < n_values; ++i) rows[i] = (int)h->rows;
    return ODEINTR_OK;
  }
  CU_TRY(cudaMemcpyAsync(rows, h->rows_inst.p, n_values * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

int odeintr_get_output_times_per_instance(odeintr_system_t h, double* t, size_t n_values, size_t inst) {
  int rc = activate(h);
  if (rc) return rc;
  if (!h->ragged) return odeintr_get_output_times(h, t, n_values);
  if (!t || inst >= h->out_m || n_values != (size_t)h->max_rows) return fail(ODEINTR_E_INVALID, "Invalid output buffer");
  CU_TRY(cudaMemcpy2DAsync(t, sizeof(double), h->t_rec.p + inst, h->out_m * sizeof(double), sizeof(double), n_values,
                           cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

int odeintr_get_stats(odeintr_system_t h, long long* accepted, long long* rejected, int* status, size_t n_inst) {
  int rc = activate(h);
  if (rc) return rc;
  const size_t have = h->sharded && h->lattice_method ? 1 : h->m;  // a slab run counts the (global) steps once
  if (n_inst != have || h->accepted.cap < have) return fail(ODEINTR_E_STATE, "no adaptive statistics recorded");
  if (accepted) CU_TRY(cudaMemcpyAsync(accepted, h->accepted.p, n_inst * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  if (rejected) CU_TRY(cudaMemcpyAsync(rejected, h->rejected.p, n_inst * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  if (status) CU_TRY(cudaMemcpyAsync(status, h->status.p, n_inst * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return ODEINTR_OK;
}

long long odeintr_last_steps(odeintr_system_t h) { return h ? h->last_steps : 0; }

// ---- device-resident access and measurement ----------------------------------------------------------
double* odeintr_state_device_ptr(odeintr_system_t h) { return h ? h->x.p : nullptr; }
double* odeintr_output_device_ptr(odeintr_system_t h) { return h ? h->out.p : nullptr; }
double* odeintr_params_device_ptr(odeintr_system_t h) { return h ? h->pars.p : nullptr; }

double odeintr_last_kernel_ms(odeintr_system_t h) {
  if (!h || !h->timed) return -1.0;
  if (cudaSetDevice(h->device) != cudaSuccess) return -1.0;
  if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}
long long odeintr_last_launches(odeintr_system_t h) { return h ? h->last_launches : 0; }
long long odeintr_total_launches(void) { return g_total_launches.load(); }

int odeintr_kernel_attributes(odeintr_system_t h, const char* kernel, int* regs, int* local_bytes, int* shared_bytes,
                              int* max_threads) {
  int rc = activate(h);
  if (rc) return rc;
  cudaKernel_t k = nullptr;
  if (!kernel || cudaLibraryGetKernel(&k, h->lib, kernel) != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(ODEINTR_E_INVALID, "no such kernel in this system");
  }
  cudaFuncAttributes fa;
  CU_TRY(cudaFuncGetAttributes(&fa, (const void*)k));
  if (regs) *regs = fa.numRegs;
  if (local_bytes) *local_bytes = (int)fa.localSizeBytes;
  if (shared_bytes) *shared_bytes = (int)fa.sharedSizeBytes;
  if (max_threads) *max_threads = fa.maxThreadsPerBlock;
  return ODEINTR_OK;
}

void* odeintr_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    g_err = "cudaHostAlloc failed";
    (void)cudaGetLastError();
    return nullptr;
  }
  return p;
}
int odeintr_host_free(void* p) {
  if (p) CU_TRY(cudaFreeHost(p));
  return ODEINTR_OK;
}

int odeintr_selftest_exact_division(int device, unsigned long long pairs, unsigned long long seed,
                                    unsigned long long* mismatches) {
  CU_TRY(cudaSetDevice(device));
  if (!mismatches) return fail(ODEINTR_E_INVALID, "null argument");
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  unsigned long long* dev = nullptr;
  CU_TRY(cudaMalloc(&dev, sizeof(unsigned long long)));
  CU_TRY(cudaMemset(dev, 0, sizeof(unsigned long long)));
  const unsigned blocks = (unsigned)prop.multiProcessorCount * 8u;
  const unsigned long long threads = (unsigned long long)blocks * 256ull;
  const unsigned long long per_thread = ((pairs + threads - 1) / threads + 7) / 8 * 8;
  odeintr_b200::launch_exact_div_check(per_thread, blocks, seed, dev, 0);
  g_total_launches += 1;
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpy(mismatches, dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaFree(dev);
  return ODEINTR_OK;
}

int odeintr_fp64_probe(int kind, int device, double* dp_instr_per_s, double* ms_out) {
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  double* sink = nullptr;
  CU_TRY(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1));
  const unsigned blocks = (unsigned)prop.multiProcessorCount * 8u;
  const int iters = 4096;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CU_TRY(cudaEventRecord(e0, 0));
    odeintr_b200::launch_fp64_probe(kind, sink, iters, blocks, 0);
    CU_TRY(cudaEventRecord(e1, 0));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
    g_total_launches += 1;
  }
  CU_TRY(cudaGetLastError());
  const double instr = (double)blocks * 256.0 * (double)iters * 64.0;  // 8 unrolled x 8 chains per iteration
  if (dp_instr_per_s) *dp_instr_per_s = instr / ((double)best * 1e-3);
  if (ms_out) *ms_out = best;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
  return ODEINTR_OK;
}

int odeintr_hbm_probe(int device, size_t bytes, double* bytes_per_s) {
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  bytes = (bytes / 16) * 16;
  if (bytes == 0) return fail(ODEINTR_E_INVALID, "probe size too small");
  void *a = nullptr, *b = nullptr;
  CU_TRY(cudaMalloc(&a, bytes)); CU_TRY(cudaMalloc(&b, bytes));
  CU_TRY(cudaMemset(a, 1, bytes));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CU_TRY(cudaEventRecord(e0, 0));
    odeintr_b200::launch_copy_probe(a, b, (long long)bytes, (unsigned)prop.multiProcessorCount * 16u, 0);
    CU_TRY(cudaEventRecord(e1, 0));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
    g_total_launches += 1;
  }
  CU_TRY(cudaGetLastError());
  if (bytes_per_s) *bytes_per_s = 2.0 * (double)bytes / ((double)best * 1e-3);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(a); cudaFree(b);
  return ODEINTR_OK;
}

int odeintr_pcie_probe(int device, void* host_in, size_t bytes_in, void* host_out, size_t bytes_out, double* seconds) {
  // Ceiling of the host-buffer path: the same pinned host ranges an odeintr_integrate_*_host call reads (host_in, H2D)
  // and writes (host_out, D2H), moved by back-to-back 256 MiB cudaMemcpyAsync copies on two streams at once -- what
  // the chunk pipeline could reach if the kernels took no time.  Either range may be null.
  CU_TRY(cudaSetDevice(device));
  if (!seconds) return fail(ODEINTR_E_INVALID, "null argument");
  const size_t chunk = (size_t)256 << 20;
  char *dev_in = nullptr, *dev_out = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e_in = nullptr;
  CU_TRY(cudaMalloc(&dev_in, chunk)); CU_TRY(cudaMalloc(&dev_out, chunk));
  CU_TRY(cudaMemset(dev_out, 0, chunk));
  CU_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CU_TRY(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1)); CU_TRY(cudaEventCreateWithFlags(&e_in, cudaEventDisableTiming));
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaEventRecord(e0, s_out));
  CU_TRY(cudaStreamWaitEvent(s_in, e0, 0));
  if (host_in)
    for (size_t o = 0; o < bytes_in; o += chunk)
      CU_TRY(cudaMemcpyAsync(dev_in, (const char*)host_in + o, std::min(chunk, bytes_in - o), cudaMemcpyHostToDevice, s_in));
  if (host_out)
    for (size_t o = 0; o < bytes_out; o += chunk)
      CU_TRY(cudaMemcpyAsync((char*)host_out + o, dev_out, std::min(chunk, bytes_out - o), cudaMemcpyDeviceToHost, s_out));
  CU_TRY(cudaEventRecord(e_in, s_in));
  CU_TRY(cudaStreamWaitEvent(s_out, e_in, 0));
  CU_TRY(cudaEventRecord(e1, s_out));
  CU_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
  *seconds = (double)ms * 1e-3;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e_in);
  cudaStreamDestroy(s_in); cudaStreamDestroy(s_out);
  cudaFree(dev_in); cudaFree(dev_out);
  return ODEINTR_OK;
}

int odeintr_device_info(int device, char* name, size_t name_cap, int* sm_count, int* cc_major, int* cc_minor,
                        size_t* total_mem) {
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (name && name_cap) { std::strncpy(name, prop.name, name_cap - 1); name[name_cap - 1] = '\0'; }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  return ODEINTR_OK;
}

}  // extern "C"
