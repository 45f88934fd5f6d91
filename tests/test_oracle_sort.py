"""The S-level SSP oracle (oracle/sort_oracle.py) and the drop-in class's initialisation against golden vectors of the
unmodified reference (CPU)."""
import os

import torch

from oracle import sort_oracle as O
from oracle import ssp_oracle as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sort_small.pt")


def _weights():
    from models import S_SSP
    return S_SSP().state_dict()


def test_dropin_initialisation_is_the_reference_stream():
    fx = torch.load(GOLD, weights_only=False)
    sd = _weights()
    assert list(sd.keys()) == fx["keys"]
    assert [tuple(v.shape) for v in sd.values()] == fx["shapes"]
    for k, c in fx["weight_checksums"].items():
        assert S.checksum(sd[k]) == c, k


def test_sort_oracle_matches_reference_golden():
    fx = torch.load(GOLD, weights_only=False)
    W = _weights()
    with torch.no_grad():
        for i, (verb, roles) in enumerate(fx["problems"]):
            trace = []
            pred, lps = O.generate_not_normal(W, verb, roles, trace=trace)
            assert pred == fx["pred"][i].tolist(), (i, pred)
            # the reference stores the log-probs of its choices in a LONG tensor (sort_model.py:118): truncated toward zero
            assert torch.equal(torch.tensor(lps).trunc().long(), fx["logp"][i]), i
            assert torch.allclose(torch.stack(trace), fx["step_rows"][i], atol=1e-5, rtol=1e-5), i
            n = sum(r != 0 for r in roles)
            assert sorted(pred[:n]) == sorted(r for r in roles if r) and all(p == 0 for p in pred[n:])


def test_prefix_states_do_not_depend_on_later_tokens():
    """What lets the device path keep a key/value cache: positions >= 1 never see <bos> (masked, weight exp(-1e3) = 0) nor later
    positions, so their states are the same in every longer prefix (sort_modules.py:126-131)."""
    W = _weights()
    with torch.no_grad():
        prior = O.encode(W, torch.tensor([[17]]), torch.tensor([[3, 9, 1, 12, 0, 0, 0, 0, 0, 0]]))
        full = O.decode(W, torch.tensor([[0, 3, 9, 1, 12]]), prior)
        for s in range(2, 5):
            part = O.decode(W, torch.tensor([[0, 3, 9, 1, 12][:s]]), prior)
            assert torch.allclose(part[:, 1:], full[:, 1:s], atol=1e-5)


def test_eval_bookkeeping_on_a_hand_case():
    import numpy as np
    # caption with 5 slots: verb 7 has roles {3 (slots 0 and 3), 5 (slot 1)}, verb 9 has roles {5 (slot 1), 2 (slot 4)}
    v = np.zeros((10, 8)); sr = np.zeros((10, 8))
    v[0, 0], sr[0, 0] = 7, 3
    v[1, 0], sr[1, 0] = 7, 5
    v[1, 1], sr[1, 1] = 9, 5
    v[3, 0], sr[3, 0] = 7, 3
    v[4, 0], sr[4, 0] = 9, 2
    roles, find, rerank = O.verb_roles(7, v, sr)
    assert roles == [3, 5] and find == {3: [0, 3], 5: [1]} and rerank == [3]
    rank = O.caption_rank([7, 9, 0], v, sr,
                          order_roles=lambda verb, roles: list(reversed(roles)) + [0],
                          order_regions=lambda role, slots: list(reversed(slots)),
                          merge=S.verb_rank_merge)
    # verb 7: roles reversed -> [5, 3] -> slots [1] + reversed([0, 3]) = [1, 3, 0]; verb 9: [2, 5] -> [4, 1]; merge inserts 4 before 1
    assert rank == [4, 1, 3, 0]
    src, verbs = O.permute_slots(rank, 10, [True, True, False, True, True] + [False] * 5, [10., 11., 12., 13., 14.] + [-1.] * 5)
    assert src == [4, 1, 3, 0] + [0] * 6
    assert verbs == [14., 11., 13., 10.] + [-1.] * 6
