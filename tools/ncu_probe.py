"""Workload for ncu captures: a few beam-search decodes of b config-2-shaped captions (eager launches when
VSRDEC_GRAPH=0, so every kernel is its own launch).  python tools/ncu_probe.py [b] [decodes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from stack_probe import inputs  # noqa: E402


def main():
    from models import ControllableCaptioningModel
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    torch.manual_seed(1234)
    m = ControllableCaptioningModel(20, 10000, 2, verb_tables=({}, {})).to("cuda:0").eval()
    stat = inputs(b, 7, "cuda:0")
    for _ in range(n):
        m.beam_search_v(stat, [3, -1], 5, 1, gt=True)
    torch.cuda.synchronize()
    print("launches per decode:", m._eng.launch_count() // n if n else 0)


if __name__ == "__main__":
    main()
