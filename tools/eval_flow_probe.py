"""Eval inner loop on synthetic data: role-ordering pre-step (vsrdec.preorder.RoleOrderer) -> slot permutation (index form) ->
beam_search_v_indexed, 100 captions per batch; serial, and with the pre-step of batch i + 1 enqueued on a side stream while
batch i decodes (order_begin / order_end).  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vsr-guided-cic_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    from models import ControllableCaptioningModel, S_SSP, SinkhornNet
    from tools.synth import synth_eval_captions
    from vsrdec.preorder import RoleOrderer, permute_slot_index
    dev = "cuda:0"
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 100          # captions per batch (pre-step call and decode call)
    K, D, R = max(8, 2000 // C), 50, 20
    torch.manual_seed(1234)
    model = ControllableCaptioningModel(20, 10000, 2, verb_tables=({}, {})).to(dev).eval()
    ro = RoleOrderer(S_SSP().to(dev).eval(), SinkhornNet(10, 20, 0.1).to(dev).eval())
    batches = []
    for i in range(4):
        d = synth_eval_captions(C=C, seed=50 + i, R=R, F=8)
        g = torch.Generator().manual_seed(70 + i)
        det = torch.relu(torch.randn((C, D, 2048), generator=g)).to(dev)
        batches.append((d, det, d["slot_index"].to(dev), d["seqs_perm"].to(dev)))

    def args(b):
        d, _, _, sp = b
        return (d["control_verb"], d["det_seqs_v"], d["det_seqs_sr"], d["verb_list"], sp, d["slot_valid"])

    def decode(b, src, verbs):
        _, det, sidx, _ = b
        return model.beam_search_v_indexed((det, permute_slot_index(sidx, src), verbs.to(dev).double()), eos_idxs=[3, -1], beam_size=5,
                                           out_size=1, gt=True)

    def serial():
        for i in range(K):
            b = batches[i % 4]
            src, verbs = ro.order(*args(b))
            out = decode(b, src, verbs)
        torch.cuda.synchronize()
        return out

    side = torch.cuda.Stream()

    def overlapped():
        with torch.cuda.stream(side):
            st = ro.order_begin(*args(batches[0]))
        for i in range(K):
            b = batches[i % 4]
            with torch.cuda.stream(side):
                src, verbs = ro.order_end(st)
                if i + 1 < K:
                    st = ro.order_begin(*args(batches[(i + 1) % 4]))
            out = decode(b, src, verbs)
        torch.cuda.synchronize()
        return out

    res = {"captions_per_batch": C, "batches": K}
    for name, fn in (("serial", serial), ("overlapped", overlapped)):
        for _ in range(2):
            fn()
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        res[name] = {"captions_per_s": K * C / dt, "ms_per_batch": dt * 1e3 / K}
    # same captions either way
    a = serial()[0][0].cpu()
    b = overlapped()[0][0].cpu()
    res["same_tokens"] = bool(torch.equal(a, b))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
