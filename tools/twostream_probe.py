"""Throughput of 1..4 independent decodes in flight (one engine and one stream each) (diagnostic)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main(b=100, k=5, D=50, L=10, R=20, V=10000, iters=20):
    from models import ControllableCaptioningModel
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    models, statics, streams = [], [], []
    for s in range(4):
        m = ControllableCaptioningModel(20, V, 2, verb_tables=({}, {})).to(dev).eval()
        g = torch.Generator(device=dev).manual_seed(1 + s)
        det = torch.relu(torch.randn((b, D, 2048), device=dev, generator=g))
        ds = torch.relu(torch.randn((b, L, R, 2048), device=dev, generator=g))
        nv = torch.randint(1, R + 1, (b, L), device=dev, generator=g)
        ds = ds * (torch.arange(R, device=dev)[None, None, :] < nv[:, :, None]).unsqueeze(-1)
        verbs = -torch.ones((b, L), dtype=torch.float64, device=dev)
        verbs[:, 2] = 17
        models.append(m); statics.append((det, ds, verbs)); streams.append(torch.cuda.Stream(dev))
    torch.cuda.synchronize()

    def run(n_streams):
        for _ in range(3):
            for s in range(n_streams):
                with torch.cuda.stream(streams[s]):
                    models[s].beam_search_v(statics[s], [3, -1], k, 1, gt=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            for s in range(n_streams):
                with torch.cuda.stream(streams[s]):
                    models[s].beam_search_v(statics[s], [3, -1], k, 1, gt=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = iters * n_streams
        print(f"{n_streams} stream(s): {1e3 * dt / n:.3f} ms per decode -> {n * b / dt:.1f} captions/s", flush=True)

    for n in (1, 2, 3, 4, 2, 3):
        run(n)


if __name__ == "__main__":
    main()
