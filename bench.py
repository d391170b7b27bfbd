#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json configs[1]).

Workload: Lorenz ensemble, 100 M random initial conditions PER GPU (weak scaling; Philox4x32-10, seed
20260101, instance index = counter, so rank r regenerates instances [r*M, (r+1)*M) of one global
ensemble), fixed-step rk4 in exact (no-FMA, Boost operation order) mode, dt = 0.001, T = 1.0 (1 000
steps), observer every 0.1 time units (every 100 steps, 11 samples per trajectory, sample-major SoA).
One "step" of the benchmark = one integrate_const pass over the rank's whole ensemble.

  value     trajectory-steps/s with the state resident in HBM (device-timed, CUDA events)
  e2e       the same metric through the host-buffer C-ABI call: initial states in pinned host memory,
            observer record and final states back in pinned host memory, copies inside the timed region
  roofline  FP64-pipe issue roofline of the rk4 kernel: 78 DP instructions per trajectory-step
            (4 x 9 RHS + 3 x 14 stage algebra) against the DP issue rate measured live by
            a DADD/DMUL probe kernel (MEASURED_PEAKS.json carries no FP64 figure)
  cpu_baseline  the oracle port of the reference's algorithm on all host cores, bounded sample

`--impl reference` times only the CPU oracle port (the reference itself cannot be built here: no R, Rcpp
or Boost in the image).
"""
import argparse
import ctypes as C
import json
import os
import pathlib
import statistics
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LORENZ = """
  dxdt[0] = sigma * (x[1] - x[0]);
  dxdt[1] = R * x[0] - x[1] - x[0] * x[2];
  dxdt[2] = -b * x[2] + x[0] * x[1];
"""
PARS = {"sigma": 10, "R": 28, "b": 8 / 3}
SEED = 20260101
LO = (-20.0, -25.0, 5.0)
HI = (20.0, 25.0, 45.0)
T_END, DT, STRIDE = 1.0, 0.001, 100
DP_INSTR_PER_STEP = 78.0      # exact mode: 4 RHS x 9 + 3 components x 14 (mul and add issue separately)
# FP64-pipe instructions per step ATTEMPT of the adaptive kernels, derived from ncu (FP64 pipe active % x cycles x 16
# lanes per scheduler and clock / attempts; profiles/r02_adaptive.md): Van der Pol dopri5 with dense output ~280 (RHS 30,
# stage algebra ~120, error norm with 2 divisions ~30, controller pow ~100), Robertson rosenbrock4 ~1000 (47 IEEE divisions
# at ~10 pipe instructions each are half of it).  With these the roofline fraction below IS the pipe utilisation.
DP_INSTR_PER_TRY_VDP = 280.0
DP_INSTR_PER_TRY_ROBERTSON = 1000.0
METRIC = "lorenz_rk4_ensemble_trajectory_steps_per_s"
UNIT = "trajectory-steps/s"


def workload_config(instances, n_gpus):
    return {
        "workload": "BASELINE.json configs[1]: Lorenz ensemble, rk4 fixed step (exact no-FMA mode, Boost operation "
                    "order), dt=0.001, T=1.0 (1000 steps), observer every 100 steps (11 samples), "
                    "x0 ~ U(-20,20)xU(-25,25)xU(5,45) Philox4x32-10 seed 20260101",
        "instances_per_gpu": instances,
        "instances_total": instances * n_gpus,
        "sharding": "instance-index ranges, one process per GPU, no data-path collective",
        "l2_policy": "inputs (2.4 GB state) and outputs (26.4 GB record) per step exceed the 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.dev) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "power_w_max": max(self.power) if self.power else None,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
def cpu_port_throughput(instances, threads, reps=1):
    """Oracle port of the reference algorithm (restated Boost odeint rk4 + integrate_const with the observer
    every step kept every 100th, as the reference's name() would record) over `instances` trajectories."""
    import numpy as np

    from oracle import build_oracle

    ora = build_oracle.build(LORENZ, PARS, True, "rk4")
    rng = np.random.default_rng(SEED)
    x0 = np.stack([rng.uniform(LO[v], HI[v], instances) for v in range(3)])
    best = None
    steps = int(round(T_END / DT))
    for _ in range(reps):
        x = x0.copy()
        t0 = time.perf_counter()
        ora.ensemble(0, x, None, 0.0, T_END, DT, obs_stride=STRIDE, out_rows=steps // STRIDE + 1, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return instances * steps / best, best


def run_reference(args, saved_stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = len(os.sched_getaffinity(0)) or 1
    # calibrate, then size each step's sample so the whole run stays within ~2 minutes
    rate, _ = cpu_port_throughput(20_000, threads)
    budget_s = 100.0 / max(1, args.steps + args.warmup)
    inst = int(min(2_000_000, max(50_000, rate * budget_s / 1000.0)))
    for _ in range(args.warmup):
        cpu_port_throughput(inst, threads)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_port_throughput(inst, threads)
        t_total += dt
    value = inst * 1000.0 * args.steps / t_total
    sample = "%d instances x 1000 rk4 steps per step (of the 100 M-instance workload), OpenMP over instances" % inst
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(100_000_000, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle port of the reference's Boost-odeint path (oracle/); the reference itself needs R, "
                "Rcpp and Boost, none of which exist in this image",
    }
    _emit(saved_stdout, line)
    return 0


# ------------------------------------------------------------------------------------------------------
_ALL_CPUS = os.sched_getaffinity(0)


def _bind_to_gpu_numa_node(index):
    """Run this rank's host threads (and hence first-touch its pinned buffers) on the CPUs NVML reports as local
    to its GPU: with 8 ranks streaming 29 GB per step each, host buffers on the wrong socket turn the end-to-end
    path into an inter-socket-link benchmark."""
    if os.environ.get("ODEINTR_BENCH_NO_BIND"):
        return
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = {w * 64 + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def _claim_stdout():
    """Keep stdout clean for the ONE JSON line: libraries (NCCL's version banner, ...) write to fd 1, so fd 1 is
    pointed at stderr for the duration of the run and the line goes to the saved descriptor at the end."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _emit(saved_fd, line):
    sys.stdout.flush()
    os.write(saved_fd, (json.dumps(line) + "\n").encode())


def main():
    saved_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=100_000_000, help="instances per GPU")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the measurements of the other BASELINE configs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args, saved_stdout)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import numpy as np
    import torch

    torch.cuda.set_device(local_rank)
    _bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import odeintr_b200
    from odeintr_b200 import _abi

    lib = _abi.load()
    assert lib.odeintr_set_device(local_rank) == 0, _abi.last_error()
    code = odeintr_b200.compile_sys("lorenz", LORENZ, PARS, True, method="rk4")
    system = code.system
    h = system._h
    M = args.instances
    steps_per_pass = 1000
    lo = (C.c_double * 3)(*LO)
    hi = (C.c_double * 3)(*HI)

    # roofline denominators measured live on this device
    dp_peak, probe_ms = C.c_double(), C.c_double()
    assert lib.odeintr_fp64_probe(1, local_rank, C.byref(dp_peak), C.byref(probe_ms)) == 0, _abi.last_error()
    dp_peak_fma = C.c_double()
    assert lib.odeintr_fp64_probe(0, local_rank, C.byref(dp_peak_fma), C.byref(probe_ms)) == 0, _abi.last_error()

    # ---- device-resident arm -------------------------------------------------------------------------
    assert lib.odeintr_fill_state_uniform(h, M, SEED, rank * M, lo, hi) == 0, _abi.last_error()

    def one_pass():
        rc = lib.odeintr_integrate_const(h, 0.0, T_END, DT, STRIDE, 1)
        if rc != 0:
            raise RuntimeError(_abi.last_error())

    for _ in range(args.warmup):
        one_pass()
    lib.odeintr_synchronize(h)
    sampler = ClockSampler(local_rank)
    launches0 = lib.odeintr_total_launches()
    barrier()
    sampler.start()
    kernel_ms = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass()
        kernel_ms.append(system.last_kernel_ms())  # CUDA events on the launching stream (syncs on the stop event)
    lib.odeintr_synchronize(h)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = lib.odeintr_total_launches() - launches0
    assert system.last_steps() == steps_per_pass and lib.odeintr_output_rows(h) == 11
    dev_s = max_over_ranks(sum(kernel_ms) * 1e-3)
    wall = max_over_ranks(wall)
    value = world * M * steps_per_pass * args.steps / dev_s
    kern_avg_ms = statistics.mean(kernel_ms)
    achieved_dp = M * steps_per_pass * DP_INSTR_PER_STEP / (kern_avg_ms * 1e-3)  # this rank's kernel
    attrs = system.kernel_attributes("odeintr_plain_ensemble")

    # ---- end-to-end arm: host buffers through odeintr_integrate_const_host --------------------------------
    e2e = None
    if not args.no_e2e:
        bytes_per_inst = (11 * 3 + 3) * 8
        try:
            avail = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) * 1024
        except Exception:
            avail = 64 << 30
        m_e2e = int(min(M, (0.55 * avail / max(world, 1)) // bytes_per_inst))
        m_e2e -= m_e2e % 128
        lib.odeintr_release_record(h)
        n_x, n_out = 3 * m_e2e, 11 * 3 * m_e2e
        px = lib.odeintr_host_alloc(n_x * 8)
        po = lib.odeintr_host_alloc(n_out * 8)
        assert px and po, "pinned allocation failed: " + _abi.last_error()
        x_host = np.ctypeslib.as_array(C.cast(px, _abi.c_dbl_p), shape=(3, m_e2e))
        out_host = np.ctypeslib.as_array(C.cast(po, _abi.c_dbl_p), shape=(11, 3, m_e2e))
        # synthetic initial states: generated once on the device (same Philox stream), copied to the host buffer
        assert lib.odeintr_fill_state_uniform(h, m_e2e, SEED, rank * M, lo, hi) == 0, _abi.last_error()
        assert lib.odeintr_get_state(h, C.cast(px, _abi.c_dbl_p), n_x) == 0, _abi.last_error()
        out_host[...] = 0.0  # touch the pages

        def e2e_pass():
            # what NAME(init, duration, step) returns is the observer record; its last row IS the final state, so no
            # separate final-state copy is requested (x_final = NULL)
            rc = lib.odeintr_integrate_const_host(h, C.cast(px, _abi.c_dbl_p), m_e2e, None, 0, 0.0, T_END, DT, STRIDE,
                                                  C.cast(po, _abi.c_dbl_p), n_out, None, 0)
            if rc != 0:
                raise RuntimeError(_abi.last_error())

        k_e2e = args.e2e_steps if args.e2e_steps else max(3, min(args.steps, 5))
        for _ in range(3):
            e2e_pass()
        barrier()
        t_e2e = 0.0
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_pass()  # returns when the record and the final state are complete in host memory
            t_e2e += system.last_kernel_ms() * 1e-3
        barrier()
        wall_e2e = max_over_ranks(time.perf_counter() - t0)
        t_e2e = max_over_ranks(t_e2e)
        # the box's ceiling for this traffic: the same pinned buffers moved by bare back-to-back copies, every rank at the
        # same moment (N ranks share the host's PCIe root complexes and memory controllers)
        ceiling = None
        if hasattr(lib, "odeintr_pcie_probe"):
            secs = C.c_double()
            t_probe = []
            for _ in range(2):
                barrier()
                if lib.odeintr_pcie_probe(local_rank, px, n_x * 8, po, n_out * 8, C.byref(secs)) != 0:
                    raise RuntimeError(_abi.last_error())
                t_probe.append(max_over_ranks(secs.value))
            ceiling = min(t_probe)
        # spot check: the record of the last pass starts at the inputs and its last row is a propagated state
        lib.odeintr_synchronize(h)
        e2e_pass()
        assert np.array_equal(out_host[0, :, :4096], x_host[:, :4096])
        assert np.all(np.isfinite(out_host[10, :, :4096])) and not np.array_equal(out_host[10, :, :4096], x_host[:, :4096])
        e2e = {
            "value": world * m_e2e * steps_per_pass * k_e2e / wall_e2e, "unit": UNIT,
            "h2d_bytes_per_step": n_x * 8, "d2h_bytes_per_step": n_out * 8,
            "instances_per_gpu": m_e2e, "steps": k_e2e, "ms_per_step": 1e3 * wall_e2e / k_e2e,
            "device_ms_per_step": 1e3 * t_e2e / k_e2e,
            "api": "odeintr_integrate_const_host (pinned host buffers, chunked 3-stream copy/compute overlap)",
        }
        if ceiling:
            e2e["pcie_ceiling"] = {
                "ms_per_step": 1e3 * ceiling, "d2h_GBps_per_gpu": n_out * 8 / ceiling / 1e9,
                "aggregate_GBps": world * (n_out + n_x) * 8 / ceiling / 1e9,
                "how": "odeintr_pcie_probe: this step's h2d + d2h bytes through the same pinned buffers as bare 256 MiB "
                       "copies on two streams, all %d rank(s) at once, device-timed, max over ranks, best of 2" % world}
            e2e["frac_of_pcie_ceiling"] = ceiling / (wall_e2e / k_e2e)
        lib.odeintr_host_free(px); lib.odeintr_host_free(po)

    # ---- secondary: the other BASELINE.json configs, reported beside the headline, never instead of it -------------
    # N = 1: configs[2] (VdP dopri5), configs[3] (Robertson rosenbrock4), configs[4] on one GPU (bistable lattice and
    # the two-field oscillator chain), the M x M lattice and the ensemble of wider systems.
    # N > 1: configs[4] proper -- the 2^28-cell lattice slab-sharded over the N GPUs with the NCCL ring halo exchange
    # (the only path with a collective), parity-asserted against the unsharded run first.
    secondary = None
    if not args.no_secondary:
        if lib.odeintr_release_record(h) != 0:
            raise RuntimeError(_abi.last_error())
        system.close()
        secondary = {}
        if world == 1:
            os.sched_setaffinity(0, _ALL_CPUS)  # the CPU samples inside the secondaries get every host core
            table = (("vdp_dopri5", vdp_secondary), ("robertson_rosenbrock4", robertson_secondary),
                     ("heterogeneous_vdp_k6", heterogeneous_secondary),
                     ("lattice_rk4", lattice_secondary), ("lattice_chain_rk4", chain_secondary),
                     ("lattice_dopri5", lattice_dopri5_secondary),
                     ("lattice2d_rk4", lattice2d_secondary), ("lattice2d_dopri5", lattice2d_dopri5_secondary),
                     ("lattice_ensemble_rk4", lattice_ensemble_secondary),
                     ("lattice_ensemble_dopri5", lattice_ensemble_adaptive_secondary))
            for key, fn in table:
                t_sec = time.perf_counter()
                try:
                    secondary[key] = fn(lib, dp_peak.value)
                except Exception as exc:  # the headline line must survive whatever happens here
                    secondary[key] = {"error": str(exc)[:300]}
                secondary[key]["bench_seconds"] = round(time.perf_counter() - t_sec, 1)
        else:
            try:
                secondary["lattice_sharded"] = lattice_sharded_secondary(lib, dp_peak.value, rank, world, local_rank, dist)
            except Exception as exc:
                secondary["lattice_sharded"] = {"error": str(exc)[:300]}

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, _ALL_CPUS)  # the CPU baseline gets every host core again
        threads = len(_ALL_CPUS)
        rate, _ = cpu_port_throughput(20_000, threads)
        inst = int(min(4_000_000, max(50_000, rate * 12.0 / 1000.0)))
        v, secs = cpu_port_throughput(inst, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d of the 100 M instances x 1000 rk4 steps, observer every 100 steps, OpenMP over instances "
                         "(%.1f s)" % (inst, secs)}

    # DRAM bytes per launch of the headline kernel: from the committed ncu --set full capture of this kernel (the spec's
    # definition of `traffic`; a run cannot count DRAM bytes itself), scaled to this run's instance count
    traffic = None
    tfile = ROOT / "profiles" / "rk4_ensemble_traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get("dram_bytes_per_launch_100M")
            if traffic is not None and M != 100_000_000:
                traffic = traffic * M / 100_000_000
        except Exception:
            traffic = None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(M, world),
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "clocks": clocks,
            "gpu_launches": int(launches),
            "e2e": e2e,
            "roofline": {
                "bound": "fp64", "achieved": achieved_dp / 1e12, "peak": dp_peak.value / 1e12, "unit": "TFLOP/s",
                "frac": achieved_dp / dp_peak.value, "traffic": traffic,
                "traffic_source": "profiles/rk4_ensemble_traffic.json (ncu --set full capture of odeintr_plain_ensemble, "
                                  "dram__bytes_read.sum + dram__bytes_write.sum; algorithmic bytes are in `hbm`)",
                "kernel": "odeintr_plain_ensemble<rk4_stepper, N=3>",
                "kernel_ms": kern_avg_ms,
                "note": "compute-bound on the FP64 pipe (north_star: FP64 pipe utilisation). achieved = 78 DP "
                        "instructions (= flops: exact mode issues no FMA) per trajectory-step x trajectory-steps per "
                        "launch / CUDA-event launch time; peak = DADD/DMUL issue rate measured live by "
                        "odeintr_fp64_probe (MEASURED_PEAKS.json has no FP64 entry); the DFMA rate is the same "
                        "%.2f T instr/s = %.1f TFLOP/s if every instruction were an FMA" %
                        (dp_peak_fma.value / 1e12, 2 * dp_peak_fma.value / 1e12),
                "hbm": {"bytes_per_launch": M * (3 + 3 + 33) * 8,
                        "achieved_GBps": M * (3 + 3 + 33) * 8 / (kern_avg_ms * 1e-3) / 1e9,
                        "peak_GBps": _measured_hbm()},
                "registers": attrs["registers"], "local_bytes": attrs["local_bytes"],
            },
            "cpu_baseline": cpu,
            "secondary": secondary,
        }
        _emit(saved_stdout, line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


BISTABLE_1D = """
for (int i = 0; i < N; ++i)
  dxdt[i] = D * (x[(i + N - 1) % N] + x[(i + 1) % N] - 2.0 * x[i]) + a * x[i] * (1 - x[i]) * (x[i] - b);
"""


def lattice_secondary(lib, dp_peak):
    """BASELINE.json configs[4] on one GPU: the reference's large-system example (R/odeintr.R:396-401) as a 1-D
    periodic lattice of 2^28 cells, fixed-step rk4 through the plain compile_sys_openmp call (temporally blocked
    kernels, stencil radius measured by the library).  50 DP instructions per cell and step (4 x 9 RHS + 14 stage
    algebra, no FMA), 16 bytes per cell and step through HBM."""
    import numpy as np

    import odeintr_b200
    from odeintr_b200 import _abi

    n, steps, dt = 1 << 28, 50, 0.01
    s = odeintr_b200.compile_sys_openmp("bistable", BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, True, method="rk4",
                                        sys_dim=n).system
    try:
        s.set_state((np.arange(n) % 7 < 3).astype(np.float64))
        ms = []
        for rep in range(4):  # the first pass is the warm-up (it also runs the probe kernel)
            if lib.odeintr_integrate_const(s._h, 0.0, steps * dt, dt, 1, 0) != 0:
                raise RuntimeError(_abi.last_error())
            if rep:
                ms.append(s.last_kernel_ms())
        t = statistics.mean(ms) * 1e-3
        rate = n * steps / t
        hbm = _measured_hbm()
        # parity + CPU sample at 2^22 cells, 10 steps: same snippet through the oracle
        ns = 1 << 22
        ss = odeintr_b200.compile_sys_openmp("bistable_s", BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, True, method="rk4",
                                             sys_dim=ns).system
        xs = (np.arange(ns) % 7 < 3).astype(np.float64)
        got = ss.no_record(xs, 10 * dt, dt)
        ss.close()
        ref, cpu = _cpu_lattice_sample(BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, ns, 10, dt, xs)
        return {
            "parity": {"bit_exact_vs_oracle_2^22_cells_10_steps": bool(np.array_equal(got.view(np.uint64), ref.view(np.uint64)))},
            "cpu_baseline": cpu,
            "workload": "BASELINE.json configs[4] on one GPU: 1-D periodic bistable lattice, 2^28 cells, rk4 fixed step, "
                        "%d steps per pass, state resident in HBM (2 GiB per array > L2)" % steps,
            "metric": "lattice_rk4_cell_steps_per_s", "value": rate, "unit": "cell-steps/s",
            "ms_per_step": 1e3 * t / steps, "passes": len(ms), "halo": s.get_halo(), "gpu_launches": s.last_launches(),
            "roofline": {"bound": "fp64", "achieved": rate * 50.0 / 1e12, "peak": dp_peak / 1e12, "unit": "TFLOP/s",
                         "frac": rate * 50.0 / dp_peak,
                         "hbm": {"bytes_per_cell_step": 8, "achieved_GBps": rate * 8.0 / 1e9, "peak_GBps": hbm,
                                 "frac": rate * 8.0 / 1e9 / hbm,
                                 "note": "two rk4 steps per pass: 8 B in + 8 B out per cell and TWO steps"},
                         "kernel": "odeintr_lattice_rk4_warp_h1",
                         "note": "the stage-per-launch schedule (odeintr_lattice_stage) moves 128 B per cell and step "
                                 "and is HBM-bound at 0.90 of the peak = 4.6e10 cell-steps/s; temporal blocking in "
                                 "registers (warp-autonomous tiles, two steps per pass) leaves 8 B and an FP64-bound step"},
        }
    finally:
        s.close()


CHAIN = """
for (int i = 0; i < L; ++i) {
  const double q = x[i];
  dxdt[i] = x[L + i];
  dxdt[L + i] = -q - beta * q * q * q + K * (x[(i + L - 1) % L] - 2.0 * q + x[(i + 1) % L]);
}
"""


def _cpu_lattice_sample(src, pars, n, steps, dt, x0, globals=""):
    """Oracle port on a bounded sample of a lattice workload: n cells x steps rk4 steps, the user's loop and the
    stepper's algebra under OpenMP on all host cores (what openmp_range_algebra does in the reference)."""
    from oracle import build_oracle

    ora = build_oracle.build(src, pars, True, "rk4", sys_dim=n, globals=globals)
    t0 = time.perf_counter()
    ref = ora.integrate_adaptive(x0, 0.0, steps * dt, dt)
    secs = time.perf_counter() - t0
    return ref["x"], {"value": n * steps / secs, "unit": "cell-steps/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                      "sample": "%d cells x %d rk4 steps (%.1f s), user loop and stepper algebra under OpenMP" % (n, steps, secs)}


def chain_secondary(lib, dp_peak):
    """SURVEY.md section 8(d) config 5 as written there: periodic chain of L = 2^27 anharmonic oscillators, SoA state
    [q, p] of 2^28 doubles, rk4, dt = 0.01, q_i(0) = sin(2 pi 8 i / L), p = 0.  Two fields; 64 DP instructions per
    site and step (4 x 9 RHS + 2 x 14 stage algebra) = 32 per state, 16 bytes per state and step through HBM."""
    import numpy as np

    import odeintr_b200
    from odeintr_b200 import _abi

    n, steps, dt = 1 << 28, 50, 0.01
    L = n // 2
    pars = {"beta": 1.0, "K": 0.5}
    glob = "static const int L = N / 2;"
    s = odeintr_b200.compile_sys_openmp("chain", CHAIN, pars, True, method="rk4", sys_dim=n, globals=glob).system
    try:
        x0 = np.zeros(n)
        x0[:L] = np.sin(2 * np.pi * 8 * np.arange(L) / L)
        s.set_state(x0)
        del x0
        ms = []
        for rep in range(4):
            if lib.odeintr_integrate_const(s._h, 0.0, steps * dt, dt, 1, 0) != 0:
                raise RuntimeError(_abi.last_error())
            if rep:
                ms.append(s.last_kernel_ms())
        t = statistics.mean(ms) * 1e-3
        rate = n * steps / t  # state-steps/s (two states per site)
        hbm = _measured_hbm()
        # parity + CPU sample at 2^21 sites, 10 steps: same snippet through the oracle
        ns = 1 << 22
        ss = odeintr_b200.compile_sys_openmp("chain_s", CHAIN, pars, True, method="rk4", sys_dim=ns, globals=glob).system
        xs = np.zeros(ns)
        xs[:ns // 2] = np.sin(2 * np.pi * 8 * np.arange(ns // 2) / (ns // 2))
        got = ss.no_record(xs, 10 * dt, dt)
        ss.close()
        ref, cpu = _cpu_lattice_sample(CHAIN, pars, ns, 10, dt, xs, glob)
        cpu["unit"] = "state-steps/s"
        return {"workload": "SURVEY 8(d) config 5: periodic chain of 2^27 anharmonic oscillators (2^28 states, fields q and "
                            "p), rk4 fixed step, %d steps per pass, state resident in HBM" % steps,
                "metric": "lattice_rk4_state_steps_per_s", "value": rate, "unit": "state-steps/s",
                "ms_per_step": 1e3 * t / steps, "halo": s.get_halo(), "gpu_launches": s.last_launches(),
                "parity": {"bit_exact_vs_oracle_2^22_states_10_steps": bool(np.array_equal(got.view(np.uint64), ref.view(np.uint64)))},
                "roofline": {"bound": "fp64", "achieved": rate * 32.0 / 1e12, "peak": dp_peak / 1e12, "unit": "TFLOP/s",
                             "frac": rate * 32.0 / dp_peak, "dp_instr_per_state_step": 32,
                             "hbm": {"bytes_per_state_step": 8, "achieved_GBps": rate * 8.0 / 1e9, "peak_GBps": hbm,
                                     "frac": rate * 8.0 / 1e9 / hbm,
                                     "note": "two rk4 steps per pass: 8 B in + 8 B out per state and TWO steps"},
                             "kernel": "odeintr_lattice_rk4_warp_h1"},
                "cpu_baseline": cpu}
    finally:
        s.close()


VDP = "dxdt[0] = x[1]; dxdt[1] = mu * (1 - x[0] * x[0]) * x[1] - x[0];"
ROBERTSON = """
dxdt[0] = -alpha * x[0] + beta * x[1] * x[2];
dxdt[1] = alpha * x[0] - beta * x[1] * x[2] - gamma * x[1] * x[1];
dxdt[2] = gamma * x[1] * x[1];
"""


def _adaptive_secondary(lib, dp_peak, name, build_system, build_oracle_fn, par_lo, par_hi, x_init, times, dt, rtol,
                        dp_instr_per_try, workload, metric):
    """Shared driver of configs[2] and configs[3]: 10 M instances with a linear parameter sweep, name_at() at the given
    times (dense output), accepted / rejected step counts, a 2 048-instance parity sample against the oracle and the
    oracle's own throughput on a bounded sample."""
    import numpy as np

    from odeintr_b200 import _abi

    m = 10_000_000
    s = build_system()
    try:
        P = len(par_lo)
        lo = (C.c_double * P)(*par_lo)
        hi = (C.c_double * P)(*par_hi)
        if lib.odeintr_fill_params_linspace(s._h, m, 0, m, lo, hi) != 0:
            raise RuntimeError(_abi.last_error())
        X = np.repeat(np.asarray(x_init, dtype=np.float64)[:, None], m, axis=1)
        tt = np.ascontiguousarray(times, dtype=np.float64)
        ms = []
        for rep in range(3):
            s.set_state(X)
            if lib.odeintr_integrate_times(s._h, tt.ctypes.data_as(_abi.c_dbl_p), tt.size, dt) != 0:
                raise RuntimeError(_abi.last_error())
            if rep:
                ms.append(s.last_kernel_ms())
        del X
        t = statistics.mean(ms) * 1e-3
        st = s.stats()
        acc, rej = int(st["accepted"].sum()), int(st["rejected"].sum())
        # parity sample: 2 048 instances spread over the sweep
        N = len(x_init)
        idx = np.linspace(0, m - 1, 2048).astype(np.int64)
        sub = np.empty((1, N, tt.size))
        got = np.empty((tt.size, N, idx.size))
        for j, i in enumerate(idx):
            if lib.odeintr_get_output(s._h, sub.ctypes.data_as(_abi.c_dbl_p), sub.size, int(i), 1) != 0:
                raise RuntimeError(_abi.last_error())
            got[:, :, j] = sub[0].T
        ora = build_oracle_fn()
        threads = len(os.sched_getaffinity(0))
        frac = idx / (m - 1)
        Ps = np.stack([par_lo[k] + (par_hi[k] - par_lo[k]) * frac for k in range(P)])
        e = ora.ensemble(1, np.repeat(np.asarray(x_init, dtype=np.float64)[:, None], idx.size, axis=1), Ps, times=tt, dt=dt,
                         out_rows=tt.size, threads=threads)
        err = float(np.max(np.abs(got - e["out"]) / (1.0 + np.abs(e["out"]))))
        # CPU sample
        mc = 100_000
        fc = np.arange(mc) / (mc - 1)
        Pc = np.stack([par_lo[k] + (par_hi[k] - par_lo[k]) * fc for k in range(P)])
        t0 = time.perf_counter()
        ec = ora.ensemble(1, np.repeat(np.asarray(x_init, dtype=np.float64)[:, None], mc, axis=1), Pc, times=tt, dt=dt,
                          out_rows=tt.size, threads=threads)
        tc = time.perf_counter() - t0
        tries = acc + rej
        attrs = s.kernel_attributes("odeintr_adaptive_ensemble")
        return {"workload": workload, "metric": metric, "value": acc / t, "unit": "accepted steps/s", "instances": m,
                "kernel_ms": 1e3 * t, "accepted": acc, "rejected": rej, "tries_per_s": tries / t,
                "samples_per_instance": int(tt.size), "record_bytes": int(tt.size * N * m * 8),
                "gpu_launches": s.last_launches(),
                "parity": {"instances_checked": int(idx.size), "max_err_over_1_plus_abs": err, "tolerance_10_rtol": 10 * rtol,
                           "ok": bool(err <= 10 * rtol),
                           "accepted_count_equal_fraction": float(np.mean(st["accepted"][idx] == e["accepted"])),
                           "rejected_count_equal_fraction": float(np.mean(st["rejected"][idx] == e["rejected"]))},
                "roofline": {"bound": "fp64", "achieved": tries / t * dp_instr_per_try / 1e12, "peak": dp_peak / 1e12,
                             "unit": "TFLOP/s", "frac": tries / t * dp_instr_per_try / dp_peak,
                             "dp_instr_per_attempt": dp_instr_per_try, "kernel": "odeintr_adaptive_ensemble (%s)" % name,
                             "registers": attrs["registers"], "local_bytes": attrs["local_bytes"],
                             "note": "FP64-pipe instructions per step attempt derived from ncu (pipe active % x cycles x "
                                     "lanes / attempts, profiles/r02_adaptive.md); unit is DP instructions, not flops"},
                "cpu_baseline": {"value": float(ec["accepted"].sum()) / tc, "unit": "accepted steps/s", "cores": threads,
                                 "kind": "port", "sample": "%d instances of the sweep (%.1f s), OpenMP over instances" % (mc, tc)}}
    finally:
        s.close()


def vdp_secondary(lib, dp_peak):
    """BASELINE.json configs[2]: Van der Pol mu-sweep, 10 M instances, mu in [0.5, 2], adaptive dopri5 with dense output
    (rk5_i, atol = rtol = 1e-8), name_at(seq(0, 20, by = 0.2)), step_size 0.01, x0 = (2, 0)."""
    import numpy as np

    import odeintr_b200
    from oracle import build_oracle

    rtol = 1e-8
    return _adaptive_secondary(
        lib, dp_peak, "dopri5_dense",
        lambda: odeintr_b200.compile_sys("vpol", VDP, "mu", method="rk5_i", atol=rtol, rtol=rtol).system,
        lambda: build_oracle.build(VDP, "mu", False, "rk5_i", atol=rtol, rtol=rtol),
        [0.5], [2.0], [2.0, 0.0], np.arange(0, 20.0001, 0.2), 0.01, rtol, DP_INSTR_PER_TRY_VDP,
        "BASELINE.json configs[2]: Van der Pol mu-sweep, 10 M instances, adaptive dopri5 (rk5_i, atol = rtol = 1e-8), "
        "dense-output name_at at 101 times, record 16 GB resident in HBM", "vdp_dopri5_dense_accepted_steps_per_s")


def robertson_secondary(lib, dp_peak):
    """BASELINE.json configs[3]: Robertson, 10 M instances (alpha in [0.02, 0.08]), compile_implicit rosenbrock4 with
    the symbolic Jacobian, name_at(10^seq(-5, 5, len = 41)), atol = rtol = 1e-6 (the reference's defaults)."""
    import numpy as np

    import odeintr_b200
    from oracle import build_oracle

    pars = {"alpha": 0.04, "beta": 1e4, "gamma": 3e7}
    return _adaptive_secondary(
        lib, dp_peak, "rosenbrock4_dense",
        lambda: odeintr_b200.compile_implicit("rob", ROBERTSON, pars).system,
        lambda: build_oracle.build(ROBERTSON, pars, False, "rosenbrock4"),
        [0.02, 1e4, 3e7], [0.08, 1e4, 3e7], [1.0, 0.0, 0.0], 10 ** np.linspace(-5, 5, 41), 1e-6, 1e-6,
        DP_INSTR_PER_TRY_ROBERTSON,
        "BASELINE.json configs[3]: Robertson alpha-sweep, 10 M instances, compile_implicit rosenbrock4 (exact mode: Boost's "
        "divisions), symbolic Jacobian, name_at at 41 log-spaced times", "robertson_rosenbrock4_accepted_steps_per_s")


def lattice_dopri5_secondary(lib, dp_peak):
    """The reference's documented large-system call on its DEFAULT method: bistable_at(inic, c(0, 1, 5, 10)) with rk5_i
    (R/odeintr.R:393-421) on a 1-D lattice of 2^26 cells: adaptive dopri5, one temporally blocked pass per attempted
    step (warp-autonomous kernel), controller on the host, dense output by repeating the accepted attempt.
    111 DP instructions per cell and attempt (6 x 9 RHS + 57 stage algebra), 32 bytes per cell and attempt."""
    import numpy as np

    import odeintr_b200
    from odeintr_b200 import _abi

    n = 1 << 26
    s = odeintr_b200.compile_sys_openmp("b1d", BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, True, sys_dim=n).system
    try:
        x0 = np.random.default_rng(1).uniform(0, 1, n)
        at = np.array([0.0, 1.0, 5.0, 10.0])
        ms = []
        for rep in range(3):
            s.set_state(x0)
            if lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) != 0:
                raise RuntimeError(_abi.last_error())
            if rep:
                ms.append(s.last_kernel_ms())
        st = s.stats()
        acc, rej = int(st["accepted"][0]), int(st["rejected"][0])
        tries = acc + rej + (at.size - 1)  # every interpolated observation repeats one attempt
        t = statistics.mean(ms) * 1e-3
        rate = n * tries / t
        # parity + CPU sample at 2^20 cells: same call through the oracle
        from oracle import build_oracle

        ns = 1 << 20
        env = {}
        ss = odeintr_b200.compile_sys_openmp("b1s", BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, True, sys_dim=ns, env=env).system
        xs = x0[:ns].copy()
        got = env["b1s_at"](xs, at, 0.1).iloc[:, 1:].to_numpy()
        sst = ss.stats()
        ss.close()
        ora = build_oracle.build(BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, True, "rk5_i", sys_dim=ns)
        t0 = time.perf_counter()
        ref = ora.integrate_times(xs, at, 0.1)
        tc = time.perf_counter() - t0
        return {"workload": "bistable_at(inic, c(0, 1, 5, 10)) with the default method rk5_i on a 1-D lattice of 2^26 cells "
                            "(adaptive dopri5, atol = rtol = 1e-6)",
                "metric": "lattice_dopri5_cell_attempts_per_s", "value": rate, "unit": "cell-attempts/s",
                "ms_per_call": 1e3 * t, "accepted": acc, "rejected": rej, "gpu_launches": s.last_launches(),
                "parity": {"bit_exact_vs_oracle_2^20_cells": bool(np.array_equal(got.view(np.uint64), ref["rec"].view(np.uint64))),
                           "step_counts_equal": bool(sst["accepted"][0] == ref["accepted"] and sst["rejected"][0] == ref["rejected"])},
                "roofline": {"bound": "fp64", "achieved": rate * 111.0 / 1e12, "peak": dp_peak / 1e12, "unit": "TFLOP/s",
                             "frac": rate * 111.0 / dp_peak, "kernel": "odeintr_lattice_dopri5_warp_h1"},
                "cpu_baseline": {"value": ns * (ref["accepted"] + ref["rejected"]) / tc, "unit": "cell-attempts/s",
                                 "cores": len(os.sched_getaffinity(0)), "kind": "port",
                                 "sample": "2^20 cells, same call (%.1f s)" % tc}}
    finally:
        s.close()


def heterogeneous_secondary(lib, dp_peak):
    """K6: a heterogeneous adaptive ensemble -- the Van der Pol sweep with mu SHUFFLED over [0.1, 10] (1 Mi instances,
    dopri5 dense output, 101 samples): neighbouring lanes hold unrelated trajectories, which reject, accept and cross
    their sample times at different moments.  Index order against odeintr_set_instance_order(COST_PROXY) (pilot + radix
    sort + index map inside the timed call); per-instance results and step counts must be identical."""
    import numpy as np

    import odeintr_b200
    from odeintr_b200 import _abi

    m = 1 << 20
    s = odeintr_b200.compile_sys("vpol", VDP, "mu", method="rk5_i", atol=1e-8, rtol=1e-8).system
    try:
        mus = np.random.default_rng(0).uniform(0.1, 10.0, m)
        s.set_params(mu=mus)
        tt = np.arange(0, 20.0001, 0.2)
        X = np.repeat(np.array([2.0, 0.0])[:, None], m, axis=1)
        res = {}
        for order in ("index", "cost"):
            s.set_instance_order(order)
            ms = []
            for rep in range(3):
                s.set_state(X)
                if lib.odeintr_integrate_times(s._h, tt.ctypes.data_as(_abi.c_dbl_p), tt.size, 0.01) != 0:
                    raise RuntimeError(_abi.last_error())
                if rep:
                    ms.append(s.last_kernel_ms())
            st = s.stats()
            res[order] = (statistics.mean(ms), st["accepted"].copy(), st["rejected"].copy(), s.get_state().copy())
        same = bool(np.array_equal(res["index"][3].view(np.uint64), res["cost"][3].view(np.uint64)) and
                    np.array_equal(res["index"][1], res["cost"][1]) and np.array_equal(res["index"][2], res["cost"][2]))
        tries = int(res["cost"][1].sum() + res["cost"][2].sum())
        return {"workload": "Van der Pol, mu shuffled over [0.1, 10], %d instances, dopri5 dense output at 101 times" % m,
                "metric": "heterogeneous_vdp_attempts_per_s", "value": tries / res["cost"][0] * 1e3, "unit": "step attempts/s",
                "ms_index_order": res["index"][0], "ms_cost_proxy_order": res["cost"][0],
                "speedup": res["index"][0] / res["cost"][0], "identical_results_and_step_counts": same,
                "roofline": {"bound": "fp64", "frac": tries / res["cost"][0] * 1e3 * DP_INSTR_PER_TRY_VDP / dp_peak,
                             "note": "step attempts x DP instructions per attempt against the measured issue peak"}}
    finally:
        s.close()


def lattice_sharded_secondary(lib, dp_peak, rank, world, local_rank, dist):
    """BASELINE.json configs[4] proper: the 1-D bistable lattice slab-sharded over the N GPUs, temporally blocked rk4
    kernels on every slab, NCCL ring halo exchange inside the C-ABI (odeintr_lattice_shard_run) every 8 steps.
    1) parity: a 1 000 003-cell lattice, 10 steps, gathered state bit-compared with rank 0's unsharded run;
    2) strong scaling: 2^28 cells over N GPUs; 3) weak scaling: 2^27 cells per GPU.  Device-timed, max over ranks."""
    import numpy as np
    import torch

    import odeintr_b200
    from odeintr_b200.distributed import ShardedLattice

    pars = {"D": 0.1, "a": 1.0, "b": 0.5}
    spe = 8

    def init(f, cells):
        return ((cells % 7) < 3).astype(np.float64)

    n_small = 1_000_003
    code = odeintr_b200.compile_sys_openmp("bs", BISTABLE_1D, pars, True, method="rk4", sys_dim=n_small)
    lat = ShardedLattice(code.system, n_small, rank=rank, world=world, steps_per_exchange=spe)
    lat.set_state_from(init)
    lat.step(10, 0.1, t0=0.0)
    full = lat.gather_state().cpu().numpy().reshape(-1)
    ok = torch.zeros(1, dtype=torch.int32, device="cuda")
    if rank == 0:
        code1 = odeintr_b200.compile_sys_openmp("bs1", BISTABLE_1D, pars, True, method="rk4", sys_dim=n_small)
        ref = code1.system.no_record(init(0, np.arange(n_small)), 1.0, 0.1)
        ok[0] = int(np.array_equal(full.view(np.uint64), ref.view(np.uint64)))
        code1.system.close()
    dist.broadcast(ok, src=0)
    parity_ok = bool(int(ok.item()))
    del lat
    code.system.close()
    out = {"workload": "BASELINE.json configs[4]: 1-D periodic bistable lattice, rk4 fixed step, slab-sharded over %d GPUs, "
                       "NCCL ring halo exchange (ncclSend/ncclRecv, ghost zone %d steps wide)" % (world, spe),
           "metric": "lattice_rk4_cell_steps_per_s", "unit": "cell-steps/s", "n_gpus": world, "steps_per_exchange": spe,
           "parity_ok": parity_ok,
           "parity": "%d cells x 10 steps sharded over %d ranks vs the unsharded single-GPU run, bit for bit" % (n_small, world)}
    if not parity_ok:
        out["error"] = "sharded result differs from the unsharded run"
        return out

    def timed(n_total, steps):
        code = odeintr_b200.compile_sys_openmp("bsl", BISTABLE_1D, pars, True, method="rk4", sys_dim=n_total)
        lat = ShardedLattice(code.system, n_total, rank=rank, world=world, steps_per_exchange=spe)
        lat.set_state_from(init)
        lat.step(2 * spe, 0.01, t0=0.0)
        torch.cuda.synchronize()
        dist.barrier()
        ms_all = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(lat.stream)
            lat.step(steps, 0.01)
            e1.record(lat.stream)
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ms_all.append(float(ms.item()))
            dist.barrier()
        # the ring exchange alone, back to back (its un-overlapped cost; the run above hides part of it)
        ex = None
        if hasattr(lib, "odeintr_lattice_exchange"):
            reps = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                lib.odeintr_lattice_exchange(code.system._h)
            torch.cuda.synchronize()
            dist.barrier()
            e0.record(lat.stream)
            for _ in range(reps):
                lib.odeintr_lattice_exchange(code.system._h)
            e1.record(lat.stream)
            torch.cuda.synchronize()
            exm = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
            dist.all_reduce(exm, op=dist.ReduceOp.MAX)
            ex = float(exm.item())
        ghost, launches = lat.ghost, int(lib.odeintr_last_launches(code.system._h))
        del lat
        code.system.close()
        best = min(ms_all)
        return {"cells_total": n_total, "cells_per_gpu": n_total // world, "steps": steps, "value": n_total * steps / best * 1e3,
                "ms_per_step": best / steps, "exchange_ms": ex, "exchange_ms_per_step": None if ex is None else ex / spe,
                "ghost_cells_per_side": ghost, "gpu_launches_per_pass": launches,
                "roofline": {"bound": "fp64", "frac": (n_total / world) * steps / best * 1e3 * 50.0 / dp_peak,
                             "note": "per GPU: 50 DP instructions per cell and step against the measured issue peak"}}

    strong = timed(1 << 28, 200)
    weak = timed((1 << 27) * world, 100)
    out.update({"value": strong["value"], "ms_per_step": strong["ms_per_step"],
                "exchange_ms_per_step": strong["exchange_ms_per_step"], "scaling": "strong", "strong": strong, "weak": weak})
    return out


BISTABLE_2D = """
for (int i = 0; i < N; ++i)
  dxdt[i] = D * laplace2D(x, i / M, i % M);
for (int i = 0; i < N; ++i)
  dxdt[i] += a * x[i] * (1 - x[i]) * (x[i] - b);
"""


def lattice2d_secondary(lib, dp_peak):
    """The reference's documented large-system example (R/odeintr.R:393-413): periodic M x M bistable lattice through
    laplace2D, M = 8192, rk4, plain compile_sys_openmp call (K4s2: warps streaming down strips of 64 columns, the four
    stages as a register pipeline).  66 DP instructions per cell and step (4 x 13 RHS + 14)."""
    import numpy as np

    import odeintr_b200
    from odeintr_b200 import _abi

    m, steps, dt = 8192, 40, 0.01
    n = m * m
    s = odeintr_b200.compile_sys_openmp("bistable2d", BISTABLE_2D, {"D": 0.1, "a": 1.0, "b": 0.5}, True, method="rk4",
                                        sys_dim=n).system
    try:
        s.set_state((np.arange(n) % 7 < 3).astype(np.float64))
        ms = []
        for rep in range(4):
            if lib.odeintr_integrate_const(s._h, 0.0, steps * dt, dt, 1, 0) != 0:
                raise RuntimeError(_abi.last_error())
            if rep:
                ms.append(s.last_kernel_ms())
        t = statistics.mean(ms) * 1e-3
        rate = n * steps / t
        launches = s.last_launches()
        # parity + CPU sample on a 2048 x 2048 lattice: the same steps through the oracle
        ms_ = 2048
        ns, steps_s = ms_ * ms_, 10
        xs = (np.arange(ns) % 7 < 3).astype(np.float64)
        env = {}
        ss = odeintr_b200.compile_sys_openmp("b2s", BISTABLE_2D, {"D": 0.1, "a": 1.0, "b": 0.5}, True, method="rk4",
                                             sys_dim=ns, env=env).system
        got = env["b2s_no_record"](xs, steps_s * dt, dt)
        ss.close()
        ref, cpu = _cpu_lattice_sample(BISTABLE_2D, {"D": 0.1, "a": 1.0, "b": 0.5}, ns, steps_s, dt, xs)
        return {"workload": "periodic 8192 x 8192 bistable lattice through laplace2D (the reference's compile_sys_openmp "
                            "example), rk4 fixed step, %d steps per pass, 512 MiB per array > L2" % steps,
                "metric": "lattice_rk4_cell_steps_per_s", "value": rate, "unit": "cell-steps/s",
                "ms_per_step": 1e3 * t / steps, "halo": s.get_halo(), "gpu_launches": launches,
                "parity": {"bit_exact_vs_oracle_2048x2048_10_steps": bool(np.array_equal(got.view(np.uint64), ref.view(np.uint64)))},
                "cpu_baseline": cpu,
                "roofline": {"bound": "fp64", "achieved": rate * 66.0 / 1e12, "peak": dp_peak / 1e12, "unit": "TFLOP/s",
                             "frac": rate * 66.0 / dp_peak, "kernel": "odeintr_lattice_rk4_stream2d"}}
    finally:
        s.close()


def lattice2d_dopri5_secondary(lib, dp_peak):
    """The reference's documented large-system example exactly as documented: bistable_at(inic, times) on the periodic
    M x M lattice with compile_sys_openmp's DEFAULT method rk5_i (R/odeintr.R:393-421), M = 8192.  K4ds2: one attempted
    dopri5 step per launch, warps streaming down strips of 64 columns, the six neighbour-dependent stages as a register
    pipeline with partial sums in place of stored stage derivatives; controller on the host.  135 DP instructions per
    cell and attempt (6 x 13 RHS + 57 stage algebra), 32 bytes per cell and attempt."""
    import numpy as np

    import odeintr_b200
    from odeintr_b200 import _abi

    m = 8192
    n = m * m
    pars = {"D": 0.1, "a": 1.0, "b": 0.5}
    s = odeintr_b200.compile_sys_openmp("b2d5", BISTABLE_2D, pars, True, sys_dim=n).system
    try:
        x0 = np.random.default_rng(1).uniform(0, 1, n)
        at = np.array([0.0, 1.0, 5.0, 10.0])
        ms = []
        for rep in range(3):
            s.set_state(x0)
            if lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) != 0:
                raise RuntimeError(_abi.last_error())
            if rep:
                ms.append(s.last_kernel_ms())
        st = s.stats()
        acc, rej = int(st["accepted"][0]), int(st["rejected"][0])
        tries = acc + rej + (at.size - 1)  # every interpolated observation repeats one attempt
        t = statistics.mean(ms) * 1e-3
        rate = n * tries / t
        launches = s.last_launches()
        # parity + CPU sample on a 1024 x 1024 lattice: the same call through the oracle
        from oracle import build_oracle

        ms_ = 1024
        ns = ms_ * ms_
        env = {}
        ss = odeintr_b200.compile_sys_openmp("b2s5", BISTABLE_2D, pars, True, sys_dim=ns, env=env).system
        xs = x0[:ns].copy()
        got = env["b2s5_at"](xs, at, 0.1).iloc[:, 1:].to_numpy()
        sst = ss.stats()
        ss.close()
        ora = build_oracle.build(BISTABLE_2D, pars, True, "rk5_i", sys_dim=ns)
        t0 = time.perf_counter()
        ref = ora.integrate_times(xs, at, 0.1)
        tc = time.perf_counter() - t0
        return {"workload": "bistable_at(inic, c(0, 1, 5, 10)) with the default method rk5_i on the periodic 8192 x 8192 lattice "
                            "through laplace2D (adaptive dopri5, atol = rtol = 1e-6)",
                "metric": "lattice_dopri5_cell_attempts_per_s", "value": rate, "unit": "cell-attempts/s",
                "ms_per_call": 1e3 * t, "accepted": acc, "rejected": rej, "gpu_launches": launches,
                "parity": {"bit_exact_vs_oracle_1024x1024": bool(np.array_equal(got.view(np.uint64), ref["rec"].view(np.uint64))),
                           "step_counts_equal": bool(sst["accepted"][0] == ref["accepted"] and sst["rejected"][0] == ref["rejected"])},
                "roofline": {"bound": "fp64", "achieved": rate * 135.0 / 1e12, "peak": dp_peak / 1e12, "unit": "TFLOP/s",
                             "frac": rate * 135.0 / dp_peak, "kernel": "odeintr_lattice_dopri5_stream2d"},
                "cpu_baseline": {"value": ns * (ref["accepted"] + ref["rejected"]) / tc, "unit": "cell-attempts/s",
                                 "cores": len(os.sched_getaffinity(0)), "kind": "port",
                                 "sample": "1024 x 1024 cells, same call (%.1f s)" % tc}}
    finally:
        s.close()


def lattice_ensemble_secondary(lib, dp_peak):
    """Ensemble of wider systems: 2^14 trajectories of a 1024-cell bistable lattice with per-trajectory parameters, rk4,
    1000 steps, observer every 100 steps; one block per trajectory, the whole call on chip (odeintr_lattice_ensemble).
    54 DP instructions per cell and step (run-time parameters: 4 x 10 RHS + 14)."""
    import numpy as np

    import odeintr_b200
    from odeintr_b200 import _abi

    cells, m, steps = 1024, 1 << 14, 1000
    s = odeintr_b200.compile_sys_openmp("bis", BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, False, method="rk4",
                                        sys_dim=cells).system
    try:
        rng = np.random.default_rng(1)
        x0 = rng.integers(0, 2, (cells, m)).astype(np.float64)
        Dv = rng.uniform(0.05, 0.3, m)
        bv = rng.uniform(0.3, 0.7, m)
        s.set_params(D=Dv, a=1.0, b=bv)
        ms = []
        for rep in range(4):
            s.set_state(x0)
            if lib.odeintr_integrate_const(s._h, 0.0, steps * 0.01, 0.01, 100, 1) != 0:
                raise RuntimeError(_abi.last_error())
            if rep:
                ms.append(s.last_kernel_ms())
        t = statistics.mean(ms) * 1e-3
        rate = cells * m * steps / t
        launches = s.last_launches()
        # parity + CPU sample: three of the trajectories through the oracle, one at a time (the reference's own loop)
        from oracle import build_oracle

        fin = s.get_state()
        ora = build_oracle.build(BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, False, "rk4", sys_dim=cells)
        ok = True
        t0 = time.perf_counter()
        picks = (0, m // 2, m - 1)
        for i in picks:
            r = ora.integrate_const(x0[:, i].copy(), 0.0, steps * 0.01, 0.01, pars=[float(Dv[i]), 1.0, float(bv[i])])
            last = np.ascontiguousarray(r["rec"][-1])
            ok = ok and bool(np.array_equal(np.ascontiguousarray(fin[:, i]).view(np.uint64), last.view(np.uint64)))
        tc = time.perf_counter() - t0
        return {"workload": "%d trajectories x %d-cell bistable lattice, per-trajectory parameters, rk4, %d steps, "
                            "observer every 100 steps" % (m, cells, steps),
                "metric": "lattice_ensemble_cell_steps_per_s", "value": rate, "unit": "cell-steps/s",
                "ms_per_pass": 1e3 * t, "gpu_launches": launches,
                "parity": {"bit_exact_vs_oracle_3_trajectories": ok},
                "cpu_baseline": {"value": cells * steps * len(picks) / tc, "unit": "cell-steps/s", "cores": 1, "kind": "port",
                                 "sample": "%d trajectories one at a time (%.2f s), single thread" % (len(picks), tc)},
                "roofline": {"bound": "fp64", "achieved": rate * 54.0 / 1e12, "peak": dp_peak / 1e12, "unit": "TFLOP/s",
                             "frac": rate * 54.0 / dp_peak, "kernel": "odeintr_lattice_ensemble"}}
    finally:
        s.close()


def lattice_ensemble_adaptive_secondary(lib, dp_peak):
    """Ensemble of wider systems under compile_sys_openmp's DEFAULT method: 2^12 trajectories of a 1024-cell bistable
    lattice with per-trajectory parameters, rk5_i (adaptive dopri5, dense output at 4 times), one block per trajectory
    with its own step sizes (K5a, lattice_ensemble_adaptive.cuh).  111 DP instructions per cell and attempt."""
    import numpy as np

    import odeintr_b200
    from odeintr_b200 import _abi

    cells, m = 1024, 1 << 12
    s = odeintr_b200.compile_sys_openmp("bisa", BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, False, sys_dim=cells).system
    try:
        rng = np.random.default_rng(2)
        x0 = rng.uniform(0, 1, (cells, m))
        Dv = rng.uniform(0.05, 0.3, m)
        bv = rng.uniform(0.3, 0.7, m)
        s.set_params(D=Dv, a=1.0, b=bv)
        at = np.array([0.0, 1.0, 5.0, 10.0])
        ms = []
        for rep in range(3):
            s.set_state(x0)
            if lib.odeintr_integrate_times(s._h, at.ctypes.data_as(_abi.c_dbl_p), at.size, 0.1) != 0:
                raise RuntimeError(_abi.last_error())
            if rep:
                ms.append(s.last_kernel_ms())
        st = s.stats()
        fin = s.get_state()
        tries = int(st["accepted"].sum() + st["rejected"].sum())
        t = statistics.mean(ms) * 1e-3
        rate = cells * tries / t
        launches = s.last_launches()
        from oracle import build_oracle

        ora = build_oracle.build(BISTABLE_1D, {"D": 0.1, "a": 1.0, "b": 0.5}, False, "rk5_i", sys_dim=cells)
        worst, counts_equal, cpu_tries = 0.0, True, 0
        t0 = time.perf_counter()
        picks = (0, m // 2, m - 1)
        for i in picks:
            r = ora.integrate_times(x0[:, i].copy(), at, 0.1, pars=[float(Dv[i]), 1.0, float(bv[i])])
            last = r["rec"][-1]
            worst = max(worst, float(np.max(np.abs(fin[:, i] - last) / (1.0 + np.abs(last)))))
            counts_equal = counts_equal and int(st["accepted"][i]) == r["accepted"] and int(st["rejected"][i]) == r["rejected"]
            cpu_tries += r["accepted"] + r["rejected"]
        tc = time.perf_counter() - t0
        return {"workload": "%d trajectories x %d-cell bistable lattice, per-trajectory parameters, default method rk5_i, "
                            "NAME_at(x0, c(0, 1, 5, 10))" % (m, cells),
                "metric": "lattice_ensemble_cell_attempts_per_s", "value": rate, "unit": "cell-attempts/s",
                "ms_per_pass": 1e3 * t, "gpu_launches": launches, "attempts": tries,
                "parity": {"max_err_over_1_plus_abs_3_trajectories": worst, "tolerance_10_rtol": 1e-5, "ok": bool(worst <= 1e-5),
                           "step_counts_equal": bool(counts_equal)},
                "roofline": {"bound": "fp64", "achieved": rate * 111.0 / 1e12, "peak": dp_peak / 1e12, "unit": "TFLOP/s",
                             "frac": rate * 111.0 / dp_peak, "kernel": "odeintr_lattice_ensemble_dopri5_dense"},
                "cpu_baseline": {"value": cells * cpu_tries / tc, "unit": "cell-attempts/s", "cores": 1, "kind": "port",
                                 "sample": "%d trajectories one at a time (%.2f s), single thread" % (len(picks), tc)}}
    finally:
        s.close()


def _measured_hbm():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    except Exception:
        return 6650.0


if __name__ == "__main__":
    sys.exit(main())
