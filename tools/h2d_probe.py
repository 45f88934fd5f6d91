"""H2D bandwidth of one 205 MB pinned buffer, with the default CPU affinity and with the process pinned to the
GPU-local NUMA node (NVML cpu affinity).  Diagnostic for bench.py's e2e figure."""
import os
import sys
import time

import torch


def bw(tag, n_bytes=205_000_000, reps=10):
    h = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda:0")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{tag}: {ms:.3f} ms per 205 MB copy = {n_bytes / ms / 1e6:.1f} GB/s, affinity={sorted(os.sched_getaffinity(0))[:4]}..({len(os.sched_getaffinity(0))} cpus)", flush=True)


def gpu_cpus(index=0):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    n = (os.cpu_count() + 63) // 64
    masks = pynvml.nvmlDeviceGetCpuAffinity(h, n)
    cpus = [64 * i + b for i, m in enumerate(masks) for b in range(64) if (m >> b) & 1]
    return cpus


if __name__ == "__main__":
    torch.cuda.init()
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    os.system("nvidia-smi topo -m 2>&1 | head -12; lscpu | grep -i numa")
    bw("default affinity")
    try:
        cpus = gpu_cpus(0)
        print("NVML affinity of GPU 0:", cpus[:8], "...", len(cpus))
        allowed = os.sched_getaffinity(0)
        use = [c for c in cpus if c in allowed]
        if use:
            os.sched_setaffinity(0, use)
            bw("gpu-local affinity")
        else:
            print("no overlap between the GPU-local cpus and the allowed set", sorted(allowed)[:8])
    except Exception as e:   # noqa: BLE001
        print("nvml affinity failed:", repr(e))
