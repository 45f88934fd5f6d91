"""Minimal workload of the eval pre-step for compute-sanitizer racecheck / synccheck (a few problems, every kernel once or twice)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from models import S_SSP, SinkhornNet  # noqa: E402
from tools.synth import synth_eval_captions  # noqa: E402
from vsrdec.preorder import RoleOrderer  # noqa: E402

dev = "cuda:0"
ro = RoleOrderer(S_SSP().to(dev).eval(), SinkhornNet(10, 20, 0.1).to(dev).eval())
d = synth_eval_captions(C=12, seed=3)
src, verbs = ro.order(d["control_verb"], d["det_seqs_v"], d["det_seqs_sr"], d["verb_list"], d["seqs_perm"].to(dev), d["slot_valid"])
torch.cuda.synchronize()
print("ordered", tuple(src.shape), int(src.sum()))
