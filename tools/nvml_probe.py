"""Which clock / throttle query disturbs a running launch stream, and by how much?  A synthetic step (300 small matmuls, ~9 ms)
is timed per step with CUDA events while the host thread -- or a sampler thread -- issues one query of each kind at known steps.
Printed: the per-step times with the query steps marked.  (bench.py's sampler choice rests on this: r2 call 11 showed single
steps of 20-150 ms whenever a poll fell into the timed region.)"""
import subprocess
import sys
import threading
import time

import pynvml
import torch

dev = torch.device("cuda:0")
a = torch.randn(1024, 1024, device=dev, dtype=torch.bfloat16)
b = torch.randn(1024, 1024, device=dev, dtype=torch.bfloat16)


def step():
    x = a
    for _ in range(300):
        x = x @ b
    return x


pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
Q = "clocks.sm,clocks.max.sm"
QR = Q + ",clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
queries = {
    "nvml clock": lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
    "nvml reasons": lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h),
    "nvml power": lambda: pynvml.nvmlDeviceGetPowerUsage(h),
    "nvml temperature": lambda: pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU),
    "smi clocks": lambda: subprocess.run(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True),
    "smi clocks+reasons": lambda: subprocess.run(["nvidia-smi", f"--query-gpu={QR}", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True),
}
for _ in range(20):
    step()
torch.cuda.synchronize()
for mode in ("main thread", "other thread"):
    for name, q in queries.items():
        n = 24
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        host = []
        marks[0].record()
        for i in range(n):
            step()
            marks[i + 1].record()
            if i in (8, 16):
                t0 = time.perf_counter()
                if mode == "main thread":
                    q()
                else:
                    th = threading.Thread(target=q)
                    th.start()
                host.append((time.perf_counter() - t0, th if mode != "main thread" else None))
        torch.cuda.synchronize()
        for _, th in host:
            if th is not None:
                th.join()
        ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(n)]
        base = sorted(ms)[n // 2]
        print(f"{mode:12s} {name:20s} median step {base:6.2f} ms, max {max(ms):7.2f} ms, steps after the queries: "
              f"{ms[9]:.2f} {ms[10]:.2f} | {ms[17]:.2f} {ms[18]:.2f}; host call {1e3 * host[0][0]:.2f} ms", flush=True)
