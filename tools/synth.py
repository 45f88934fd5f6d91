"""Synthetic COCO/Flickr30k-Entities-shaped decoder inputs (SURVEY.md §8d recipe).  Data only —
no model arithmetic — shared by the tests, the oracle-side golden generator and bench.py."""
from typing import Dict, List, Optional, Sequence, Tuple

import torch


def synth_inputs(b: int, D: int, L: int, R: int, Fd: int, seed: int, vocab_size: int,
                 n_det_range: Tuple[int, int] = None, real_slots: Tuple[int, int] = (4, 8),
                 verb_slots: Sequence[int] = (2,), verb_vocab_id: Optional[int] = 17,
                 verb_id_range: Optional[Tuple[int, int]] = None, one_region_slots: bool = False):
    """Synthetic COCO/Flickr30k-Entities-shaped decoder inputs (SURVEY.md §8d recipe).

    Features are relu(randn) (non-negative like Faster-RCNN pool5 features, so an all-zero
    row is exactly a padding row); each slot has ``n_valid ~ U{1..R}`` non-zero rows; verb
    slots hold one row = mean of the image's valid detections (data/field.py:460,517);
    slots after the last real one repeat it and carry verb -1 (eval_coco.py:231-237).
    Returns det (b,D,F) f32, det_seqs (b,L,R,F) f32, verbs (b,L) f64.
    """
    g = torch.Generator().manual_seed(seed)
    det = torch.relu(torch.randn((b, D, Fd), generator=g))
    if n_det_range is not None:
        n_det = torch.randint(n_det_range[0], n_det_range[1] + 1, (b,), generator=g)
        for i in range(b):
            det[i, int(n_det[i]):] = 0
    det_seqs = torch.zeros((b, L, R, Fd))
    verbs = -torch.ones((b, L), dtype=torch.float64)
    n_real = torch.randint(real_slots[0], min(real_slots[1], L) + 1, (b,), generator=g)
    for i in range(b):
        nr = int(n_real[i])
        valid_rows = det[i][det[i].sum(-1) != 0]
        for l in range(nr):
            if l in verb_slots:
                det_seqs[i, l, 0] = valid_rows.mean(0)
                if verb_id_range is not None:
                    verbs[i, l] = float(torch.randint(verb_id_range[0], verb_id_range[1], (1,), generator=g))
                else:
                    verbs[i, l] = float(verb_vocab_id if verb_vocab_id is not None else 0)
            else:
                nv = 1 if one_region_slots else int(torch.randint(1, R + 1, (1,), generator=g))
                pick = torch.randint(0, valid_rows.size(0), (nv,), generator=g)
                det_seqs[i, l, :nv] = valid_rows[pick]
        det_seqs[i, nr:] = det_seqs[i, nr - 1]
    return det, det_seqs, verbs


def synth_verb_table(n_verbs: int, vocab_size: int, seed: int) -> Dict[str, List[int]]:
    """Synthetic verb -> vocabulary-forms table in the JSON shape the reference loads
    (controllable_captioning.py:25-34): {str(verb_id): [vocab idx, ...]}; some verbs have an
    empty list and some ids are absent to exercise the fallback (:291-292)."""
    g = torch.Generator().manual_seed(seed)
    table = {}
    for v in range(n_verbs):
        r = int(torch.randint(0, 10, (1,), generator=g))
        if r == 0:
            continue            # missing key
        n = 0 if r == 1 else int(torch.randint(1, 7, (1,), generator=g))
        table[str(v)] = [int(x) for x in torch.randint(0, vocab_size, (n,), generator=g)]
    return table


def synth_inputs_indexed(b: int, D: int, L: int, R: int, Fd: int, seed: int, n_det_range: Tuple[int, int] = None,
                         real_slots: Tuple[int, int] = (4, 8), verb_slots: Sequence[int] = (2,),
                         verb_vocab_id: Optional[int] = 17):
    """Index form of synth_inputs: det (b,D,F) f32, slot_index (b,L,R) int32 (>= 0 detection row, -2 image mean
    row, -1 padding), verbs (b,L) f64.  Slots after the last real one repeat it, as in eval_coco.py:231-237."""
    g = torch.Generator().manual_seed(seed)
    det = torch.relu(torch.randn((b, D, Fd), generator=g))
    n_det = torch.full((b,), D, dtype=torch.long)
    if n_det_range is not None:
        n_det = torch.randint(n_det_range[0], n_det_range[1] + 1, (b,), generator=g)
        for i in range(b):
            det[i, int(n_det[i]):] = 0
    idx = -torch.ones((b, L, R), dtype=torch.int32)
    verbs = -torch.ones((b, L), dtype=torch.float64)
    n_real = torch.randint(real_slots[0], min(real_slots[1], L) + 1, (b,), generator=g)
    for i in range(b):
        nr = int(n_real[i])
        for l in range(nr):
            if l in verb_slots:
                idx[i, l, 0] = -2
                verbs[i, l] = float(verb_vocab_id)
            else:
                nv = int(torch.randint(1, R + 1, (1,), generator=g))
                idx[i, l, :nv] = torch.randint(0, int(n_det[i]), (nv,), generator=g).to(torch.int32)
        idx[i, nr:] = idx[i, nr - 1]
    return det, idx, verbs


def materialize_slots(det: torch.Tensor, slot_index: torch.Tensor) -> torch.Tensor:
    """det_seqs (b,L,R,F) that the reference's fields would have built for these indices
    (data/field.py:461-541): copies of detection rows, the mean of the valid detections for -2, zeros for -1."""
    b, L, R = slot_index.shape
    Fd = det.size(2)
    out = torch.zeros((b, L, R, Fd), dtype=det.dtype)
    for i in range(b):
        valid = det[i][det[i].sum(-1) != 0]
        mean = valid.mean(0) if valid.size(0) > 0 else torch.zeros(Fd)
        sel = slot_index[i].long()
        rows = det[i][sel.clamp(min=0)]
        rows = torch.where((sel >= 0).unsqueeze(-1), rows, torch.zeros_like(rows))
        rows = torch.where((sel == -2).unsqueeze(-1), mean.expand_as(rows), rows)
        out[i] = rows
    return out


def synth_eval_captions(C=40, seed=21, R=4, F=32):
    """Synthetic inputs of the eval pre-step (coco_scripts/eval_coco.py:127-147): per caption 1-3 control verbs, 3-8 filled slots,
    each slot carrying 1-2 (verb, role) pairs — with repeated roles, so that the R-level network is needed — a verb list, the
    (vis, txt, pos) rows, small slot tiles and their index form."""
    import numpy as np
    import random
    rnd = random.Random(seed)
    g = torch.Generator().manual_seed(seed)
    control_verb = np.zeros((C, 8), dtype=np.int64)
    det_seqs_v = np.zeros((C, 10, 8), dtype=np.int64)
    det_seqs_sr = np.zeros((C, 10, 8), dtype=np.int64)
    verb_list = -np.ones((C, 10, 1), dtype=np.float64)
    slot_valid = np.zeros((C, 10), dtype=bool)
    for c in range(C):
        verbs = rnd.sample(range(1, 2663), rnd.randint(1, 3))
        control_verb[c, :len(verbs)] = verbs
        n_slots = rnd.randint(3, 8)
        slot_valid[c, :n_slots] = True
        role_pool = {v: rnd.sample(range(1, 26), rnd.randint(2, 4)) for v in verbs}
        for j in range(n_slots):
            pairs = rnd.sample(verbs, min(len(verbs), rnd.randint(1, 2)))
            for k, v in enumerate(pairs):
                det_seqs_v[c, j, k] = v
                det_seqs_sr[c, j, k] = rnd.choice(role_pool[v])
            if rnd.random() < 0.25:
                verb_list[c, j, 0] = rnd.choice(verbs)
    seqs_perm = torch.relu(torch.randn((C, 10, 2352), generator=g))
    seqs_perm[:, :, 2348:] = torch.rand((C, 10, 4), generator=g)
    tiles = torch.relu(torch.randn((C, 10, R, F), generator=g)) + 0.1
    tiles = tiles * torch.from_numpy(slot_valid).view(C, 10, 1, 1)
    slot_index = torch.randint(0, 50, (C, 10, R), generator=g, dtype=torch.int32)
    slot_index = torch.where(torch.from_numpy(slot_valid).view(C, 10, 1), slot_index, torch.full_like(slot_index, -1))
    return dict(control_verb=control_verb, det_seqs_v=det_seqs_v, det_seqs_sr=det_seqs_sr, verb_list=verb_list, slot_valid=slot_valid,
                seqs_perm=seqs_perm, tiles=tiles, slot_index=slot_index)
