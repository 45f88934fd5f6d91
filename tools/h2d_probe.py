"""Aggregate pinned host->device bandwidth of one node: every rank (one per GPU under torchrun) copies a 205 MB pinned
buffer (one bench batch) to its GPU at the same time; reports per-rank and aggregate GB/s, with the default CPU
affinity and with the process (and its pinned allocation) bound to the GPU-local CPUs (NVML cpu affinity).
The ceiling for bench.py's `e2e` figure with the reference's input format (2 MB of fp32 slot tiles per caption).

    python tools/h2d_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py"""
import json
import os

import torch


def gpu_cpus(index):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    n = (os.cpu_count() + 63) // 64
    masks = pynvml.nvmlDeviceGetCpuAffinity(h, n)
    return [64 * i + b for i, m in enumerate(masks) for b in range(64) if (m >> b) & 1]


def measure(dev, barrier, n_bytes=204_808_000, reps=12, chunks=1):
    h = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)                                   # first touch under the current affinity -> NUMA placement
    d = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    step = (n_bytes + chunks - 1) // chunks

    def copy():
        for o in range(0, n_bytes, step):
            d[o:o + step].copy_(h[o:o + step], non_blocking=True)
    for _ in range(2):
        copy()
    torch.cuda.synchronize(dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        copy()
    e1.record()
    torch.cuda.synchronize(dev)
    barrier()
    return n_bytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("gloo")          # host-side barrier / gather only: no device traffic besides the copies

    def barrier():
        if world > 1:
            dist.barrier()

    def gather(x):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    res = {}
    res["default_affinity"] = gather(measure(dev, barrier))
    res["default_affinity_4_chunks"] = gather(measure(dev, barrier, chunks=4))
    note = None
    try:
        gi = int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local]) if "CUDA_VISIBLE_DEVICES" in os.environ else local
        cpus = gpu_cpus(gi)
        use = [c for c in cpus if c in os.sched_getaffinity(0)]
        if use:
            os.sched_setaffinity(0, use)
            res["gpu_local_affinity"] = gather(measure(dev, barrier))
            note = f"rank {rank}: {len(use)} GPU-local cpus of {os.cpu_count()}"
        else:
            note = "no overlap between the GPU-local cpus and the allowed set"
    except Exception as e:   # noqa: BLE001
        note = "nvml affinity failed: " + repr(e)
    notes = gather(note)
    if rank == 0:
        out = {"n_gpus": world, "bytes_per_copy": 204_808_000, "cpu_count": os.cpu_count(), "affinity_notes": notes[:2]}
        for k, v in res.items():
            out[k] = {"per_rank_gbs": [round(x, 1) for x in v], "aggregate_gbs": round(sum(v), 1)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
