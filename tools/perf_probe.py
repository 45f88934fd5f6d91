"""Quick device-side timing probe (not the bench contract): config-2-shaped decode with per-phase
CUDA-event times from libvsrdec's profiler."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main(b=100, k=5, D=50, L=10, R=20, V=10000, iters=5):
    from models import ControllableCaptioningModel
    torch.manual_seed(0)
    m = ControllableCaptioningModel(20, V, 2, verb_tables=({}, {})).to("cuda:0").eval()
    g = torch.Generator(device="cuda:0").manual_seed(1)
    det = torch.relu(torch.randn((b, D, 2048), device="cuda:0", generator=g))
    ds = torch.relu(torch.randn((b, L, R, 2048), device="cuda:0", generator=g))
    nv = torch.randint(1, R + 1, (b, L), device="cuda:0", generator=g)
    ds = ds * (torch.arange(R, device="cuda:0")[None, None, :] < nv[:, :, None]).unsqueeze(-1)
    verbs = -torch.ones((b, L), dtype=torch.float64, device="cuda:0")
    verbs[:, 2] = 17
    statics = (det, ds, verbs)
    for _ in range(3):
        m.beam_search_v(statics, [3, -1], k, 1, gt=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        m.beam_search_v(statics, [3, -1], k, 1, gt=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"decode b={b} k={k}: {ms:.3f} ms -> {b / ms * 1e3:.1f} captions/s")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.beam_search_v(statics, [3, -1], k, 1, gt=True)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue time of one decode: {1e3 * (t1 - t0):.3f} ms (then {1e3 * (t2 - t1):.3f} ms until the GPU drains)")
    m._eng.set_profiling(True)
    m.beam_search_v(statics, [3, -1], k, 1, gt=True)
    tot = 0.0
    for name, t, n in m._eng.phase_times():
        print(f"  {name:28s} {t:9.3f} ms  ({n} calls)")
        tot += t
    print(f"  sum of phases {tot:.3f} ms")
    m._eng.set_profiling(False)


if __name__ == "__main__":
    main()
