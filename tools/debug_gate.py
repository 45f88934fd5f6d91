"""Locate the worst gate log-prob discrepancies of the device path vs the oracle on config 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from oracle import vsr_oracle as O
from common import load_golden
from gpu_common import make_model, device_beam

fx = load_golden("full_cfg1.pt")
d = O.Dims()
W = O.init_weights(d, seed=1234)
m = make_model(d, W)
print("gemm kind:", m._engine().gemm_kind())
det, ds, verbs = O.synth_inputs(**fx["synth"])
dev = tuple(t.to("cuda:0") for t in (det, ds, verbs))
(w, g), (lw, lg), hist, extra = device_beam(m, dev, [3, -1], 3, 1, True, True)
parent, word, gate, score = [x.cpu() for x in hist]
T = parent.shape[0]
forced = O.BeamTrace([parent[t].long() for t in range(T)], [word[t].long() for t in range(T)],
                     [gate[t].long() for t in range(T)], [], [], [], [])
rows = []
def hook(t, out, gate_lp, flat):
    n = gate_lp.size(0)
    dg = extra["step_gate"][t, :n].cpu()
    err = (dg - gate_lp).abs()
    for r in range(n):
        for c in range(2):
            rows.append((float(err[r, c]), t, r, c, float(dg[r, c]), float(gate_lp[r, c])))
with torch.no_grad():
    O.beam_search(W, d, (det, ds, verbs), [3, -1], 3, 3, use_verbs=True, gt=True, forced=forced, step_hook=hook)
rows.sort(reverse=True)
for e, t, r, c, dv, ov in rows[:12]:
    print(f"abs_err={e:.3e} t={t} row={r} col={c} device={dv:.7f} oracle={ov:.7f}")
