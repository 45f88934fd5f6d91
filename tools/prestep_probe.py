"""Where the eval pre-step's time goes (vsrdec.preorder.RoleOrderer on 100 synthetic captions): host bookkeeping, S-level
device call, R-level device call, host assembly.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    from models import S_SSP, SinkhornNet
    from tools.synth import synth_eval_captions
    from vsrdec import preorder as P
    dev = "cuda:0"
    sort_net, sk = S_SSP().to(dev).eval(), SinkhornNet(10, 20, 0.1).to(dev).eval()
    ro = P.RoleOrderer(sort_net, sk)
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    d = synth_eval_captions(C=C, seed=5)
    sp = d["seqs_perm"].to(dev)
    args = (d["control_verb"], d["det_seqs_v"], d["det_seqs_sr"], d["verb_list"], sp, d["slot_valid"])
    for _ in range(3):
        ro.order(*args)
    torch.cuda.synchronize()
    # whole call
    t0 = time.perf_counter()
    for _ in range(10):
        ro.order(*args)
    torch.cuda.synchronize()
    whole = (time.perf_counter() - t0) * 100
    # pieces: wrap the two device calls
    acc = {"s_level_ms": 0.0, "r_level_ms": 0.0}
    g0, a0 = sort_net.generate_batch, sk.assign

    def timed(fn, key):
        def w(*a, **k):
            torch.cuda.synchronize(); t = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize(); acc[key] += (time.perf_counter() - t) * 1e3
            return r
        return w
    sort_net.generate_batch, sk.assign = timed(g0, "s_level_ms"), timed(a0, "r_level_ms")
    t0 = time.perf_counter()
    for _ in range(10):
        ro.order(*args)
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 100
    out = {"captions": C, "ms_per_call": whole, "with_syncs_ms": total, "s_level_ms": acc["s_level_ms"] / 10, "r_level_ms": acc["r_level_ms"] / 10}
    out["host_ms"] = out["with_syncs_ms"] - out["s_level_ms"] - out["r_level_ms"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
