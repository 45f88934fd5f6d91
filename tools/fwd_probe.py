"""Timing of the small-batch drivers (teacher-forced forward, greedy, sampling; 100 rows per step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from models import ControllableCaptioningModel
torch.manual_seed(1234)
dev = "cuda:0"
m = ControllableCaptioningModel(20, 10000, 2, verb_tables=({}, {})).to(dev).eval()
g = torch.Generator().manual_seed(1005)
det = torch.relu(torch.randn((100, 100, 2048), generator=g)).to(dev)
caps = torch.randint(0, 10000, (100, 20), generator=g).to(dev)
ctrl = torch.relu(torch.randn((100, 20, 20, 2048), generator=g))
nv = torch.randint(1, 21, (100, 20), generator=g)
ctrl = (ctrl * (torch.arange(20)[None, None, :] < nv[:, :, None]).unsqueeze(-1)).to(dev)
ds = ctrl[:, :10].contiguous()


def timeit(fn, n=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("forward_teacher ms:", round(timeit(lambda: m((det,), (caps, ctrl))), 3))
print("greedy ms:", round(timeit(lambda: m.test(det, ds)), 3))
print("sample_rl ms:", round(timeit(lambda: m.sample_rl(det, ds, seed=1)), 3))

# per-phase breakdown of one forward (eager, CUDA events around every phase)
eng = m._eng
eng.set_profiling(True)
for _ in range(3):
    m((det,), (caps, ctrl))
torch.cuda.synchronize()
print("forward phases (3 eager forwards):", eng.phase_times())
eng.set_profiling(False)
