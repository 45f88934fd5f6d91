"""Golden vectors of the S-level SSP step from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_sort.py     ->  tests/golden/sort_small.pt

Imports /root/reference/models/sort_model.py (S_SSP), builds it with its own seeded initialisation, runs
`generate(this_verb, det_seqs_sr, mode='not-normal')` (the call of coco_scripts/eval_coco.py:174) on seeded (verb, role set)
problems and stores problems + outputs + a checksum of every state_dict entry: the pin for oracle/sort_oracle.py and for the
drop-in class's initialisation (the weights themselves are 152 MB; they are reproducible from the seed)."""
import importlib.util
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VSR_REFERENCE_ROOT", "/root/reference")


def problems(n=32, seed=11):
    rnd = random.Random(seed)
    out = []
    for i in range(n):
        k = 1 + i % 10 if i < 20 else rnd.randint(2, 6)
        roles = rnd.sample(range(1, 26), k)
        out.append((rnd.randint(1, 2662), roles + [0] * (10 - k)))
    return out


def main():
    sys.path.insert(0, REF)
    from models.sort_model import S_SSP            # the reference's package (REF is first on sys.path)
    sys.path.insert(0, ROOT)
    from oracle import ssp_oracle as S
    net = S_SSP().eval()
    probs = problems()
    preds, lps, rows, seen = [], [], [], []
    # the reference keeps seqLogprobs in a LONG tensor (sort_model.py:118: det_seqs_sr.new_zeros), i.e. truncated; the float rows
    # are taken from a forward hook on the 512 -> 26 projection instead
    net.expander_nn.register_forward_hook(lambda m, i, o: seen.append(torch.log_softmax(o.detach(), -1)[0].clone()))
    with torch.no_grad():
        for verb, roles in probs:
            del seen[:]
            pred, lp, _ = net.generate(torch.tensor([verb]), torch.tensor([roles]), mode='not-normal')
            preds.append(pred[0].clone())
            lps.append(lp[0].clone())
            rows.append(torch.stack(seen))
    sd = net.state_dict()
    torch.save({"problems": probs, "pred": torch.stack(preds), "logp": torch.stack(lps), "step_rows": rows,
                "keys": list(sd.keys()), "shapes": [tuple(v.shape) for v in sd.values()],
                "weight_checksums": {k: S.checksum(v) for k, v in sd.items()}, "seed_w": 1234},
               os.path.join(HERE, "sort_small.pt"))
    # the drop-in class must reproduce the reference's initialisation bit for bit
    spec = importlib.util.spec_from_file_location("_dropin_sort_model", os.path.join(ROOT, "vsr-guided-cic_b200", "models", "sort_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mine = mod.S_SSP().state_dict()
    assert list(mine.keys()) == list(sd.keys()), "state_dict keys differ"
    bad = [k for k in sd if not torch.equal(sd[k], mine[k])]
    assert not bad, f"initialisation differs: {bad[:5]}"
    print("wrote sort_small.pt:", len(probs), "problems; drop-in init == reference init for", len(sd), "tensors")
    print(torch.stack(preds)[:12])


if __name__ == "__main__":
    main()
