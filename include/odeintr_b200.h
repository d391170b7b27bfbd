/* include/odeintr_b200.h — C-ABI of the B200-native odeintr integration engine.
 *
 * This is the drop-in boundary for the reference's hot path: everything the reference does inside
 * the translation unit it generates from inst/templates/compile_sys_template.cpp (and the
 * _openmp / _implicit variants) and hands to Rcpp::sourceCpp (R/odeintr.R:340) happens behind these
 * entry points instead.  The host generator (R or Python) emits a CUDA translation unit holding the
 * user's snippet as a __device__ functor; odeintr_compile() NVRTC-compiles it for sm_100a together
 * with the hand-written stepper kernels and loads it.  The remaining functions are what the
 * reference's generated Rcpp exports call into Boost odeint for, generalised from one trajectory to
 * an ensemble of m independent instances (initial-condition batches / parameter sweeps).
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every host buffer, the library owns every
 *    device buffer until odeintr_destroy().
 *  - every function returns 0 on success and a non-zero ODEINTR_E_* code on failure;
 *    odeintr_last_error() returns the message for the calling thread (what the R glue passes to
 *    Rcpp::stop()).
 *  - ensemble arrays are structure-of-arrays with the instance index fastest:
 *        state  x[v * m + i]          v < sys_dim, i < m
 *        params p[k * m + i]          k < n_pars
 *        output out[(s * sys_dim + v) * m + i]   s < rows
 *    With m == 1 these degenerate to the reference's single-trajectory vectors and the output to the
 *    data-frame columns X1..XN of name_get_output() (compile_sys_template.cpp:58-73).
 *  - there is no CPU fallback: every compute entry point fails with ODEINTR_E_CUDA when no
 *    sm_100-class device is usable.
 */
#ifndef ODEINTR_B200_H
#define ODEINTR_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct odeintr_system_s* odeintr_system_t;

enum {
  ODEINTR_OK = 0,
  ODEINTR_E_INVALID = 1,   /* bad argument ("Invalid initial state", "Invalid parameter vector", ...) */
  ODEINTR_E_COMPILE = 2,   /* NVRTC rejected the generated translation unit (see odeintr_compile_log) */
  ODEINTR_E_CUDA = 3,      /* CUDA runtime / driver error, or no usable device */
  ODEINTR_E_STEP = 4,      /* at least one instance hit Boost's 500-failed-steps overflow */
  ODEINTR_E_STATE = 5      /* call sequence error (e.g. continue_at with an empty record) */
};

/* compile flags */
#define ODEINTR_FLAG_FMAD 1u            /* allow FMA contraction (fast mode; NOT bit-compatible with the reference) */
#define ODEINTR_FLAG_RECIPROCAL 32u      /* rosenbrock4: divide by the step size and the LU pivots through reciprocals computed once per attempt (fast mode; agrees with the reference to rounding, NOT bit for bit) */
#define ODEINTR_FLAG_KEEP_ZERO_TERMS 2u /* also add Boost's zero tableau entries (x + 0*dt*k) */
#define ODEINTR_FLAG_LATTICE_EULER 8u   /* large-system TU stepped with euler (1 stage) instead of rk4 (4 stages): set by the generator */
#define ODEINTR_FLAG_LATTICE_ADAPTIVE 16u /* large-system TU stepped with adaptive dopri5 (rk5_a; rk5_i with DENSE_OUTPUT): set by the generator */
#define ODEINTR_FLAG_LATTICE_CK54 64u   /* large-system TU stepped with Cash-Karp 5(4) (rk54; rk54_a with LATTICE_ADAPTIVE): set by the generator */
#define ODEINTR_FLAG_LATTICE_RKF78 128u /* large-system TU stepped with Fehlberg 7(8) (rk78; rk78_a with LATTICE_ADAPTIVE): set by the generator */
#define ODEINTR_FLAG_LATTICE_RK5 256u   /* large-system TU stepped with dopri5 as a plain fixed-step stepper (rk5): set by the generator */
#define ODEINTR_FLAG_LATTICE_BS 512u    /* large-system TU stepped with Bulirsch-Stoer (bs with LATTICE_ADAPTIVE; bsd with DENSE_OUTPUT too): set by the generator */
#define ODEINTR_FLAG_DENSE_OUTPUT 4u    /* the TU instantiates a dense-output stepper (rk5_i, rosenbrock4): set by the generator */

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Replaces Rcpp::sourceCpp(code = code, env = env) at R/odeintr.R:340,457,602.
 * cuda_src  : translation unit emitted by the generator (templates/compile_*_template.cu)
 * sys_dim   : N  (__SYS_SIZE__, compile_sys_template.cpp:17)
 * n_pars    : number of run-time parameters (0 when pars are const / absent)
 * atol,rtol : error tolerances baked into the stepper constructor by R/utils.R:22-32 and
 *             compile_implicit_template.cpp:30-33
 */
int odeintr_compile(const char* cuda_src, const char* name, int sys_dim, int n_pars, double atol, double rtol,
                    unsigned flags, odeintr_system_t* out);

/* Compile only (no device needed): NVRTC -> sm_100a cubin.  *cubin_size receives the size; when
 * cubin is non-null and cubin_cap is large enough the image is copied out.  Used by the CPU-side
 * build check and tests. */
int odeintr_compile_cubin(const char* cuda_src, unsigned flags, void* cubin, size_t cubin_cap, size_t* cubin_size);

int odeintr_destroy(odeintr_system_t h);
const char* odeintr_last_error(void);
const char* odeintr_compile_log(odeintr_system_t h);
/* Select the CUDA device (ordinal) that later odeintr_compile() calls on this thread bind to;
 * one process per GPU calls this once with its local rank. */
int odeintr_set_device(int device);
/* Use an existing CUDA stream (cudaStream_t / CUstream) for all work of this system; 0 = own stream. */
int odeintr_set_stream(odeintr_system_t h, void* cuda_stream);
int odeintr_synchronize(odeintr_system_t h);
/* Per-instance budget of step attempts (accepted + rejected) in one integrate call with an adaptive stepper
 * (default 20 000 000).  Boost has no such limit and can loop forever, e.g. when the step size underflows
 * at a finite-time blow-up; on a GPU that would hang the device, so the kernels stop the instance, set
 * ODEINTR_STATUS_STEP_BUDGET and the call returns ODEINTR_E_STEP. */
int odeintr_set_max_tries(odeintr_system_t h, long long max_tries);

/* Order in which the instances of an ensemble are dealt to the GPU's lanes by the adaptive steppers (controlled and
 * dense-output: rk5_a, rk5_i, rk54_a, rk78_a, bs, bsd, rosenbrock4).  Results never depend on it: every trajectory is
 * integrated by its own lane with its own step sizes and lands at its own index.  What changes is which trajectories
 * share a warp: 32 lanes execute one instruction stream, so a lane whose neighbours reject, accept and cross their
 * sample times at other moments pays for their branches too.  In the reference this loop is serial
 * (README.md:157-163: one NAME_set_params() + NAME() per instance), so the question does not arise there.
 *   ODEINTR_ORDER_INDEX        lane j integrates instance j (default; right for sorted parameter sweeps)
 *   ODEINTR_ORDER_COST_PROXY   every integrate call first runs a pilot -- the first 12 step attempts of every instance
 *                              on a scratch copy -- and sorts the instances by the step size the controller settled on
 *                              (stable radix sort; index map perm[lane] = instance).  Shuffled or random ensembles
 *                              regain the coherence of a sorted sweep.
 *   ODEINTR_ORDER_PERMUTATION  the caller's own order: perm[0..n) (a permutation of 0..n-1, host memory, copied)
 * perm / n are ignored for the first two modes.  odeintr_get_instance_order returns the map used by the last call. */
#define ODEINTR_ORDER_INDEX 0
#define ODEINTR_ORDER_COST_PROXY 1
#define ODEINTR_ORDER_PERMUTATION 2
int odeintr_set_instance_order(odeintr_system_t h, int mode, const int* perm, size_t n);
int odeintr_get_instance_order(odeintr_system_t h, int* perm, size_t n);

/* ---- state and parameters ------------------------------------------------------------------- */

/* name_set_state(new_state), compile_sys_template.cpp:75-83.  x holds sys_dim x n_inst doubles (SoA).
 * n_inst instances are resident afterwards.  Fails with "Invalid initial state" when n_values is not
 * sys_dim * n_inst. */
int odeintr_set_state(odeintr_system_t h, const double* x, size_t n_values, size_t n_inst);
/* name_get_state(), compile_sys_template.cpp:86-90.  Copies sys_dim x n_inst doubles. */
int odeintr_get_state(odeintr_system_t h, double* x, size_t n_values);
/* name_set_params(...), pars_setter_template.cpp:1-5 / pars_vec_setter_template.cpp:1-7.
 * p holds n_pars x n_inst doubles; n_inst == 1 broadcasts one parameter set to every instance.
 * Fails with "Invalid parameter vector" on a size mismatch. */
int odeintr_set_params(odeintr_system_t h, const double* p, size_t n_values, size_t n_inst);
/* name_get_params(), pars_getter_template.cpp:1-7. */
int odeintr_get_params(odeintr_system_t h, double* p, size_t n_values);
size_t odeintr_n_instances(odeintr_system_t h);
size_t odeintr_n_param_sets(odeintr_system_t h);

/* Synthetic ensembles generated on the device: x[v][i] = lo[v] + (hi[v]-lo[v]) * u, u from
 * Philox4x32-10 with key = seed and counter = (first_instance + i, v / 2); instance i of a shard gets
 * the same numbers whatever the sharding. */
int odeintr_fill_state_uniform(odeintr_system_t h, size_t n_inst, unsigned long long seed,
                               unsigned long long first_instance, const double* lo, const double* hi);
/* p[k][i] = lo[k] + (hi[k]-lo[k]) * (first_instance + i) / (total_instances - 1)  (parameter sweep) */
int odeintr_fill_params_linspace(odeintr_system_t h, size_t n_inst, unsigned long long first_instance,
                                 unsigned long long total_instances, const double* lo, const double* hi);

/* ---- integration (the hot path) ------------------------------------------------------------- */

/* odeint::integrate_const(stepper, sys, state, t0, t1, dt, obs) — compile_sys_template.cpp:154
 * (record != 0) and :121,:134 (record == 0, the advance-to-first-time call of name_at).
 * obs_stride > 1 is the ensemble extension "observe every obs_stride-th step" for plain steppers
 * (1 = the reference's behaviour). */
int odeintr_integrate_const(odeintr_system_t h, double t0, double t1, double dt, long long obs_stride, int record);
/* odeint::integrate_times(stepper, sys, state, times.begin(), times.end(), dt, obs) — :123,:136 */
int odeintr_integrate_times(odeintr_system_t h, const double* times, size_t n_times, double dt);
/* odeint::integrate_adaptive(stepper, sys, state, t0, t1, dt[, obs]) — :107 (record) and :168 */
int odeintr_integrate_adaptive(odeintr_system_t h, double t0, double t1, double dt, int record);

/* Host-buffer form of odeintr_integrate_const: x (sys_dim x n_inst), pars (n_pars x par_sets, par_sets in
 * {1, n_inst}; null when the system has no run-time parameters) and the results live in caller memory
 * (pinned memory makes the copies asynchronous).  The ensemble is streamed through the GPU in chunks of
 * `chunk` instances (0 = automatic) with copies and kernels overlapped; out (rows x sys_dim x n_inst,
 * null = no observer) and x_final (sys_dim x n_inst, may be null) are complete when the call returns.
 * Nothing stays resident on the device.  rows = odeintr_const_rows(t0, t1, dt, obs_stride). */
int odeintr_integrate_const_host(odeintr_system_t h, const double* x, size_t n_inst, const double* pars,
                                 size_t par_sets, double t0, double t1, double dt, long long obs_stride,
                                 double* out, size_t out_values, double* x_final, size_t chunk);
/* rows integrate_const records for (t0, t1, dt, obs_stride) with a fixed-step stepper */
size_t odeintr_const_rows(double t0, double t1, double dt, long long obs_stride);

/* Temporally blocked rk4 step for 1-D compile_sys_openmp systems.  When the loop body of a cell reads x at most
 * `halo` cells to the left / right of the cell (periodic distance, in every field), a fixed-step rk4 step runs as
 * ONE kernel pass: tile + 4*halo cells per side staged in shared memory, all four stages on chip, 16 B instead of
 * 128 B of HBM traffic per cell and step; bit-identical results.
 *   halo < 0   measure the radius (a probe kernel evaluates every cell's index expressions at the next integrate
 *              call) and use the tiled kernels when it is at most 32 cells; `fields` is ignored.  Every access of
 *              every step is still checked: if a read ever reaches further (indices that depend on the state or on
 *              time), the integrate call is repeated with the stage-per-launch kernels.
 *   halo > 0   declared radius; a violation makes the integrate call fail with ODEINTR_E_INVALID.
 *   halo = 0   stage-per-launch kernels (the state a freshly compiled system is in).
 * odeintr_lattice_get_halo returns the radius in use, or -1 when the stage-per-launch kernels are. */
int odeintr_lattice_set_halo(odeintr_system_t h, long long halo, int fields);
long long odeintr_lattice_get_halo(odeintr_system_t h);

/* ---- large systems sharded over several GPUs (one process per GPU) ------------------------------------
 * The state of a compile_sys_openmp system is `fields` arrays of `cells_total` cells (sys_dim = fields *
 * cells_total; entry (f, c) is x[f * cells_total + c]).  A rank owns the cells [first_cell, first_cell +
 * n_local) of every field.  Its local buffer x_dev (device memory owned by the caller, e.g. a torch tensor, so
 * the caller may also exchange halos itself) is laid out [field][pitch] with the owned cells of a field at
 * [pad, pad + n_local) and the ghost cells directly around them: [pad - ghost, pad) and [pad + n_local,
 * pad + n_local + ghost), ghost = halo * stages (odeintr_lattice_ghost_width).  pad = ghost rounded up to 16
 * doubles and pitch = pad + n_local + pad rounded up to 16 (odeintr_lattice_pad_width / _pitch) keep local
 * cell 0 of every field 128-byte aligned.  One halo exchange per step is enough because every Runge-Kutta stage
 * recomputes a window that shrinks by the stencil radius `halo`.
 * odeintr_lattice_shard_step() advances the owned cells by one step of size dt from time t; the ghost cells of
 * x_dev must hold the neighbours' current values on entry (periodic ring) and are stale on return. */
int odeintr_lattice_shard(odeintr_system_t h, long long cells_total, long long first_cell, long long n_local,
                          int fields, int halo, double* x_dev);
long long odeintr_lattice_ghost_width(odeintr_system_t h, int halo);
/* rk4 slabs with a stencil radius of at most 32 cells step through the temporally blocked kernels (see
 * odeintr_lattice_set_halo): one pass per step, 16 B per cell.  Their ghost zone can be made `steps` times wider
 * (ghost = 4 * halo * steps), so that odeintr_lattice_shard_run exchanges halos only every `steps` steps: each step
 * consumes 4 * halo ghost cells per side and recomputes the rest (communication-avoiding, bit-identical).  Call
 * before odeintr_lattice_ghost_width / _pad_width / _pitch and odeintr_lattice_shard.  Default 1. */
int odeintr_lattice_set_exchange_interval(odeintr_system_t h, int steps);
long long odeintr_lattice_pad_width(odeintr_system_t h, int halo);
long long odeintr_lattice_pitch(odeintr_system_t h, int halo, long long n_local);
int odeintr_lattice_shard_step(odeintr_system_t h, double t, double dt);
/* Native multi-GPU loop: the library owns an NCCL communicator over the ranks' GPUs (NVLink / NVSwitch).
 * Rank 0 creates a 128-byte unique id (odeintr_nccl_unique_id) and hands it to every rank out of band
 * (torch.distributed broadcast, MPI, a file ...); every rank then calls odeintr_lattice_comm_init (collective).
 * odeintr_lattice_shard_run() does n_steps of { ring halo exchange (ncclSend / ncclRecv to both neighbours in one
 * group, on the system's stream), rk4 step at t0 + (first_step + s)*dt } without returning to the host language.
 * world == 1 needs no id: the slab's ghost cells are filled from its own far ends. */
int odeintr_nccl_unique_id(void* id_out, size_t cap);
int odeintr_lattice_comm_init(odeintr_system_t h, int world, int rank, const void* id_bytes, size_t id_size);
int odeintr_lattice_shard_run(odeintr_system_t h, long long first_step, long long n_steps, double t0, double dt);
/* One ring halo exchange of the slab registered with odeintr_lattice_shard, on the system's stream: the ghost cells
 * of x_dev receive the ring neighbours' boundary cells (what odeintr_lattice_shard_run does between rounds).  For
 * callers that step with odeintr_lattice_shard_step, and for timing the exchange alone.  Collective over the ranks. */
int odeintr_lattice_exchange(odeintr_system_t h);
/* Adaptive steppers on a sharded large system (compile_sys_openmp's default method rk5_i, and rk5_a): after
 * odeintr_lattice_shard the ordinary integrate calls -- odeintr_integrate_const / _times / _adaptive -- run on this
 * rank's slab.  They are COLLECTIVE: every rank calls them with the same arguments.  An attempted step is one
 * temporally blocked pass per slab (6 * halo ghost cells of x and of its derivative on either side, refreshed by the
 * ring exchange after every accepted step) and one 16-byte ncclAllReduce(max) of the error norm, so all ranks take
 * the same decisions with the controller the single-GPU run uses: the result is bit-identical to it.
 * The observer records whole slabs; odeintr_lattice_shard_get_output returns the owned cells of every recorded row:
 * out[(row * fields + f) * n_local + c], rows = odeintr_output_rows(h), Time column from odeintr_get_output_times,
 * step counts (global, identical on all ranks) from odeintr_get_stats(h, ..., 1). */
int odeintr_lattice_shard_get_output(odeintr_system_t h, double* out, size_t n_values);

/* ---- observer record ------------------------------------------------------------------------ */

/* name_reset_observer(), compile_sys_template.cpp:92-97 */
int odeintr_reset_observer(odeintr_system_t h);
/* free the device memory of the observer record and scratch buffers (the record is dropped) */
int odeintr_release_record(odeintr_system_t h);
/* rows currently recorded (length of rec_t) */
size_t odeintr_output_rows(odeintr_system_t h);
/* rec_t: the Time column (instance independent for const/times integration) */
int odeintr_get_output_times(odeintr_system_t h, double* t, size_t n_values);
/* name_get_output() for instances [inst_begin, inst_begin + inst_count): x receives, per instance,
 * sys_dim columns of `rows` doubles:  x[((j * sys_dim) + v) * rows + s]. */
int odeintr_get_output(odeintr_system_t h, double* x, size_t n_values, size_t inst_begin, size_t inst_count);
/* the whole record in device layout out[(s * sys_dim + v) * m + i] (rows * sys_dim * m doubles) */
int odeintr_get_output_soa(odeintr_system_t h, double* out, size_t n_values);
/* 1 when the record is ragged: integrate_adaptive with a controlled / dense stepper observes once per
 * accepted step, so every instance has its own row count and Time column; odeintr_output_rows() is then
 * the maximum over instances and shorter instances are padded (unspecified values) */
int odeintr_output_is_ragged(odeintr_system_t h);
/* per-instance row counts / times for ragged records (integrate_adaptive with adaptive steppers) */
int odeintr_get_output_rows_per_instance(odeintr_system_t h, int* rows, size_t n_values);
int odeintr_get_output_times_per_instance(odeintr_system_t h, double* t, size_t n_values, size_t inst);

/* accepted / rejected step counts and ODEINTR_STATUS_* per instance (adaptive steppers) */
int odeintr_get_stats(odeintr_system_t h, long long* accepted, long long* rejected, int* status, size_t n_inst);
/* number of fixed steps the last plain-stepper call executed per instance */
long long odeintr_last_steps(odeintr_system_t h);

/* ---- device-resident access and measurement -------------------------------------------------- */

double* odeintr_state_device_ptr(odeintr_system_t h);   /* sys_dim x m */
double* odeintr_output_device_ptr(odeintr_system_t h);  /* rows x sys_dim x m */
double* odeintr_params_device_ptr(odeintr_system_t h);
/* device time (CUDA events on the system's stream) and kernel launches of the last integrate call */
double odeintr_last_kernel_ms(odeintr_system_t h);
long long odeintr_last_launches(odeintr_system_t h);
long long odeintr_total_launches(void);
int odeintr_kernel_attributes(odeintr_system_t h, const char* kernel, int* regs, int* local_bytes,
                              int* shared_bytes, int* max_threads);

/* pinned host memory for end-to-end transfers */
void* odeintr_host_alloc(size_t bytes);
int odeintr_host_free(void* p);

/* Self-test of the correctly rounded repeated-divisor division the exact-mode kernels use (kernels/exact_div.cuh:
 * rosenbrock4's divisions by the step size and the LU pivots, dopri5's dense-output constants): `pairs` quotients
 * a / b -- random and adversarial significands, exponents inside and outside the guarded window -- are computed both
 * ways on the device and compared bit for bit; *mismatches receives the count (0 = identical). */
int odeintr_selftest_exact_division(int device, unsigned long long pairs, unsigned long long seed,
                                    unsigned long long* mismatches);

/* FP64-pipe issue-rate probe (roofline denominator for the compute-bound ensembles): returns DP
 * instructions per second summed over the device; kind 0 = DFMA, 1 = DADD/DMUL mix (no FMA). */
int odeintr_fp64_probe(int kind, int device, double* dp_instr_per_s, double* ms);
/* device-to-device copy bandwidth probe in bytes (read + write) per second */
int odeintr_hbm_probe(int device, size_t bytes, double* bytes_per_s);
/* Ceiling of the host-buffer path (odeintr_integrate_const_host): moves the pinned host range host_in to the device
 * and a device buffer to the pinned host range host_out, both at once, in back-to-back 256 MiB copies, and reports the
 * device-timed duration.  With every rank of a box calling it at the same moment this is what the host's PCIe / memory
 * system lets N GPUs move concurrently -- the denominator of bench.py's e2e.frac_of_pcie_ceiling. */
int odeintr_pcie_probe(int device, void* host_in, size_t bytes_in, void* host_out, size_t bytes_out, double* seconds);
int odeintr_device_info(int device, char* name, size_t name_cap, int* sm_count, int* cc_major, int* cc_minor,
                        size_t* total_mem);

#ifdef __cplusplus
}
#endif
#endif /* ODEINTR_B200_H */
