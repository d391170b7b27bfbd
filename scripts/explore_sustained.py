"""Sustained vs burst HBM bandwidth on a B200 + clocks while the lattice kernel chain runs."""
import ctypes as C, os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import odeintr_b200
from odeintr_b200 import _abi
from systems import BISTABLE_1D, BISTABLE_PARS
import pynvml
pynvml.nvmlInit()
dev = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []
stop = threading.Event()
def sampler():
    while not stop.is_set():
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_MEM),
                        pynvml.nvmlDeviceGetPowerUsage(dev) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(dev),
                        pynvml.nvmlDeviceGetTemperature(dev, 0)))
        stop.wait(0.02)
th = threading.Thread(target=sampler, daemon=True); th.start()
def mark(label, t0, t1):
    sel = [s for s in samples if t0 <= s[0] <= t1]
    if sel:
        print("  %s: sm %d-%d MHz, mem %d-%d MHz, power %.0f-%.0f W, reasons %s, temp %d-%d" % (label, min(s[1] for s in sel), max(s[1] for s in sel), min(s[2] for s in sel), max(s[2] for s in sel),
              min(s[3] for s in sel), max(s[3] for s in sel), sorted(set(hex(s[4]) for s in sel)), min(s[5] for s in sel), max(s[5] for s in sel)), flush=True)

# (a) torch copy: burst vs sustained
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
for reps in (1, 10, 100, 400):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(reps): b.copy_(a)
    e1.record(); torch.cuda.synchronize(); t1 = time.time()
    ms = e0.elapsed_time(e1)
    print("torch copy x%d: %.1f GB/s" % (reps, reps * 2 * n * 2 / ms / 1e6), flush=True)
    mark("copy x%d" % reps, t0, t1)
del a, b
torch.cuda.empty_cache()

lib = _abi.load()
n = 1 << 28
code = odeintr_b200.compile_sys_openmp("bistable", BISTABLE_1D, BISTABLE_PARS, True, method="rk4", sys_dim=n)
s = code.system
s.set_state((np.arange(n) % 7 < 3).astype(np.float64))
for steps in (5, 20, 100, 100):
    t0 = time.time()
    rc = lib.odeintr_integrate_const(s._h, 0.0, steps * 0.01, 0.01, 1, 0)
    ms = s.last_kernel_ms(); t1 = time.time()
    print("lattice steps=%d(%d) %.2f ms -> %.2f ms/step, %.1f GB/s" % (steps, s.last_steps(), ms, ms / s.last_steps(), 128.0 * n * s.last_steps() / ms / 1e6), flush=True)
    mark("lattice %d" % steps, t0, t1)
stop.set()
