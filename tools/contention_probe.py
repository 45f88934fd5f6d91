"""Does a concurrent 205 MB pinned H2D copy slow the decode down?  (diagnostic for bench.py's e2e figure)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main(b=100, k=5, D=50, L=10, R=20, V=10000, iters=10):
    from models import ControllableCaptioningModel
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    m = ControllableCaptioningModel(20, V, 2, verb_tables=({}, {})).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1)
    det = torch.relu(torch.randn((b, D, 2048), device=dev, generator=g))
    ds = torch.relu(torch.randn((b, L, R, 2048), device=dev, generator=g))
    nv = torch.randint(1, R + 1, (b, L), device=dev, generator=g)
    ds = ds * (torch.arange(R, device=dev)[None, None, :] < nv[:, :, None]).unsqueeze(-1)
    verbs = -torch.ones((b, L), dtype=torch.float64, device=dev)
    verbs[:, 2] = 17
    statics = (det, ds, verbs)
    host = torch.empty(205_000_000, dtype=torch.uint8).pin_memory()
    sink = torch.empty(205_000_000, dtype=torch.uint8, device=dev)
    cs = torch.cuda.Stream(dev)

    def run(copy, sync_each):
        for _ in range(3):
            m.beam_search_v(statics, [3, -1], k, 1, gt=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(iters):
            if copy:
                with torch.cuda.stream(cs):
                    sink.copy_(host, non_blocking=True)
            out, _ = m.beam_search_v(statics, [3, -1], k, 1, gt=True)
            if sync_each:
                out[0].cpu()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / iters * 1e3
        print(f"copy={copy} sync_each={sync_each}: {e0.elapsed_time(e1) / iters:.3f} ms per decode (device events), {wall:.3f} ms wall", flush=True)

    run(False, False)
    run(True, False)

    # the same decode replayed from a CUDA graph (one launch instead of ~140)
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for _ in range(3):
            m.beam_search_v(statics, [3, -1], k, 1, gt=True)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        gout, _ = m.beam_search_v(statics, [3, -1], k, 1, gt=True)
    torch.cuda.synchronize()
    ref, _ = m.beam_search_v(statics, [3, -1], k, 1, gt=True)
    graph.replay()
    torch.cuda.synchronize()
    print("graph replay equals eager:", bool(torch.equal(ref[0], gout[0])))
    for copy in (False, True):
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            if copy:
                with torch.cuda.stream(cs):
                    sink.copy_(host, non_blocking=True)
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"graph replay, copy={copy}: {e0.elapsed_time(e1) / iters:.3f} ms per decode", flush=True)


if __name__ == "__main__":
    main()
