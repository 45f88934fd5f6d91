"""CPU oracle for the S-level SSP step of the eval pre-step — TEST INFRASTRUCTURE ONLY.

Restates, as plain functions over an explicit state_dict (torch CPU ops in the reference's order):
  * TransformerEmbedding.forward        /root/reference/models/transformer_modules.py:188-212  (embedding * sqrt(d); no positional term:
                                         S_SSP() is built with pos_enc=False and the decoder never asks for one)
  * MultiHeadAttention / KeyValAttention /root/reference/models/transformer_modules.py:36-54, 104-134 (8 heads, logits / sqrt(head_dim),
                                         masked_fill(mask == 0, -1e3), softmax)
  * TransformerEncoderLayer             /root/reference/models/transformer_modules.py:325-346  (pre-LN, residuals)
  * TransformerEncoder.forward          /root/reference/models/sort_modules.py:52-63           (verb + role embeddings -> fc_feat -> 3 layers -> LN)
  * TransformerDecoderLayer.forward     /root/reference/models/sort_modules.py:79-99           (the SAME attention module serves the self and
                                         the cross attention; cross_attention's weights are never used)
  * TransformerDecoder.forward          /root/reference/models/sort_modules.py:120-135         (key j visible to query i iff j <= i and token_j != 0)
  * S_SSP.generate(mode='not-normal')   /root/reference/models/sort_model.py:105-119, 149-183  (greedy over the roles still to be placed)
  * the role bookkeeping of the eval loop /root/reference/coco_scripts/eval_coco.py:148-237    (roles of a verb, repeated roles, rank
                                         assembly, merge over verbs, permutation + tail padding of the slot list)
Only tests/ may import this module; the product path (vsr-guided-cic_b200/) never does.

Pin: tests/golden/sort_small.pt is generated from the UNMODIFIED reference class imported from /root/reference
(tests/golden/make_golden_sort.py: seeded S_SSP(), `generate(mode='not-normal')` per problem); tests/test_oracle_sort.py checks
this oracle against it (orders exactly, log-probs to 1e-5) and checks that the drop-in class's seeded initialisation is the
reference's bit for bit (per-tensor checksums).  The eval-loop restatement has no callable counterpart in the reference (it is the
body of a script that needs the COCO annotations): it is pinned piecewise — S_SSP.generate, SinkhornNet, verb_rank_merge —
and by construction properties in the tests."""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

N_HEAD = 8
N_LAYERS = 3
MAX_LEN = 10


def _lin(W, name, x):
    return F.linear(x, W[name + ".weight"], W[name + ".bias"])


def _ln(W, name, x):
    return F.layer_norm(x, (x.shape[-1],), W[name + ".weight"], W[name + ".bias"], 1e-5)


def mha(W, prefix: str, query, keys, values, mask=None):
    """transformer_modules.py:104-134 + 36-54.  query (B, Tq, d); keys/values (B, Tk, d); mask (B, 1, Tq, Tk) bool or None."""
    B, d = query.shape[0], query.shape[-1]
    hd = d // N_HEAD
    q = _lin(W, prefix + ".linear_Q", query).view(B, -1, N_HEAD, hd).transpose(1, 2)
    k = _lin(W, prefix + ".linear_K", keys).view(B, -1, N_HEAD, hd).transpose(1, 2)
    v = _lin(W, prefix + ".linear_V", values).view(B, -1, N_HEAD, hd).transpose(1, 2)
    logits = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd)
    if mask is not None:
        logits = logits.masked_fill(mask == 0, -1e3)
    ctx = torch.matmul(F.softmax(logits, dim=-1), v)
    ctx = ctx.transpose(1, 2).contiguous().view(B, -1, d)
    return _lin(W, prefix + ".linear_O", ctx)


def _ff(W, prefix, x):
    return _lin(W, prefix + ".w_2", F.relu(_lin(W, prefix + ".w_1", x)))


def encode(W, verb: torch.Tensor, roles: torch.Tensor) -> torch.Tensor:
    """sort_modules.py:52-63.  verb (B, 1) long, roles (B, L) long (0 = padding, embedded like any other id) -> (B, L, d)."""
    d = W["sr_embed_layer.weight"].shape[1]
    x = F.embedding(verb, W["v_embed_layer.weight"]) * math.sqrt(d) + F.embedding(roles, W["sr_embed_layer.weight"]) * math.sqrt(d)
    if "encoder.fc_feat.weight" in W:
        x = _lin(W, "encoder.fc_feat", x)
    for l in range(N_LAYERS):
        p = f"encoder.encoder_layers.{l}"
        y1 = _ln(W, p + ".layer_norm1", x)
        y1 = mha(W, p + ".attention", y1, y1, y1) + x
        x = _ff(W, p + ".ff_layer", _ln(W, p + ".layer_norm2", y1)) + y1
    return _ln(W, "encoder.layer_norm", x)


def decode(W, tokens: torch.Tensor, prior: torch.Tensor) -> torch.Tensor:
    """sort_modules.py:120-135, 79-99.  tokens (B, s) long (position 0 = <bos> = 0), prior (B, L, d) -> states (B, s, d)."""
    B, s = tokens.shape
    d = W["sr_embed_layer.weight"].shape[1]
    length_mask = (tokens == 0).unsqueeze(1).float()
    x = F.embedding(tokens, W["sr_embed_layer.weight"]) * math.sqrt(d)
    self_mask = torch.triu(torch.ones((s, s)), diagonal=1).unsqueeze(0)
    self_mask = ((self_mask + length_mask).unsqueeze(1) == 0)
    for l in range(N_LAYERS):
        p = f"decoder.encoder_layers.{l}"
        h1 = _ln(W, p + ".layer_norm1", x)
        h1 = mha(W, p + ".attention", h1, h1, h1, self_mask) + x
        h2 = _ln(W, p + ".layer_norm2", h1)
        h2 = mha(W, p + ".attention", h2, prior, prior) + h1          # the self-attention module again (sort_modules.py:88)
        x = _ff(W, p + ".ff_layer", _ln(W, p + ".layer_norm3", h2)) + h2
    return _ln(W, "decoder.layer_norm", x)


def step_logprobs(W, tokens: torch.Tensor, prior: torch.Tensor) -> torch.Tensor:
    """sort_model.py:158-160: log-softmax over the 26 role ids of the last position."""
    return F.log_softmax(_lin(W, "expander_nn", decode(W, tokens, prior)[:, -1]), dim=-1)


def generate_not_normal(W, verb: int, roles: Sequence[int], trace: Optional[list] = None) -> Tuple[List[int], List[float]]:
    """sort_model.py:105-119, 149-183 for ONE problem: `roles` is the zero-padded (max_len) list of distinct role ids of the verb.
    Returns (pred, seq_logprobs), both of length max_len, zero after the last placed role.  `trace`, if given, receives the
    (26,) log-prob row of every step (for tie-aware comparisons)."""
    roles = list(roles) + [0] * (MAX_LEN - len(roles))
    v = torch.tensor([[int(verb) % 10000]], dtype=torch.long)
    r = torch.tensor([roles], dtype=torch.long)
    remain = [x != 0 for x in roles]
    prior = encode(W, v, r)
    pred, lps = [0] * MAX_LEN, [0.0] * MAX_LEN
    tokens = [0]
    for t in range(MAX_LEN + 1):
        if not any(remain):
            break
        lp = step_logprobs(W, torch.tensor([tokens], dtype=torch.long), prior)[0]
        if trace is not None:
            trace.append(lp.clone())
        cand = [i for i, m in enumerate(remain) if m]
        sub = lp[torch.tensor([roles[i] for i in cand])]
        best = int(torch.max(sub, -1)[1])
        pos = cand[best]
        remain[pos] = False
        pred[t], lps[t] = roles[pos], float(sub[best])
        tokens.append(roles[pos])
    return pred, lps


# ------------------------------------------------------------------ eval-loop bookkeeping (eval_coco.py:148-237)
def verb_roles(verb: int, det_seqs_v: np.ndarray, det_seqs_sr: np.ndarray) -> Tuple[List[int], Dict[int, List[int]], List[int]]:
    """eval_coco.py:152-169: the distinct roles of `verb` in first-seen order (at most 10), the slots holding each role, and the
    roles held by more than one slot (in first-repeat order; the reference keeps them in a set)."""
    roles, find, rerank = [], {}, []
    for j in range(det_seqs_v.shape[0]):
        for k in range(det_seqs_v.shape[1]):
            if det_seqs_v[j][k] == verb and len(roles) < 10:
                sr = int(det_seqs_sr[j][k])
                if sr not in find:
                    find[sr] = [j]
                    roles.append(sr)
                else:
                    find[sr].append(j)
                    if sr not in rerank:
                        rerank.append(sr)
    return roles, find, rerank


def caption_rank(control_verb, det_seqs_v, det_seqs_sr, order_roles, order_regions, merge) -> List[int]:
    """eval_coco.py:148-215 for one caption.  order_roles(verb, roles) -> the S-level order of the role ids;
    order_regions(role, slots) -> the R-level order of the slots holding a repeated role; merge = verb_rank_merge."""
    ranks = []
    for verb in control_verb:
        verb = int(verb)
        if verb == 0:
            break
        roles, find, rerank = verb_roles(verb, det_seqs_v, det_seqs_sr)
        if not roles:
            continue
        sr_rank = {sr: order_regions(sr, find[sr]) for sr in rerank}
        rank = []
        for sr in order_roles(verb, roles):
            if sr == 0:
                break
            rank += list(sr_rank[sr]) if len(find[sr]) != 1 else find[sr]
        ranks.append(rank)
    if not ranks:
        return []
    final = ranks[0]
    for nxt in ranks[1:]:
        final = merge(final, nxt)
    return [int(x) for x in final]


def permute_slots(final_rank: Sequence[int], n_slots: int, slot_valid: Sequence[bool], verb_list: Sequence[float]):
    """eval_coco.py:217-237 on slot indices instead of slot tiles: row j of the re-ordered caption is slot final_rank[j]; rows
    whose tile is empty are dropped; the tail repeats the last kept slot; verb ids follow their slots, -1 where no slot was placed.
    Returns (src_slot (n_slots,) int, -1 = empty row; verbs (n_slots,))."""
    placed = [int(r) for r in list(final_rank)[:n_slots]]
    kept = [r for r in placed if slot_valid[r]]
    src = [-1] * n_slots
    for j, r in enumerate(kept):
        src[j] = r
    if kept:
        for j in range(len(kept), n_slots):
            src[j] = kept[-1]
    verbs = [-1.0] * n_slots
    for j, r in enumerate(placed):
        verbs[j] = float(verb_list[r])
    return src, verbs


def replay(W, verb: int, roles: Sequence[int], chosen: Sequence[int]) -> List[torch.Tensor]:
    """Log-prob rows of every step along a GIVEN order `chosen` (the device's): what a tie-aware comparison needs — after a
    near-tie resolved the other way the free-running order is a different, equally valid trajectory."""
    roles = list(roles) + [0] * (MAX_LEN - len(roles))
    prior = encode(W, torch.tensor([[int(verb) % 10000]], dtype=torch.long), torch.tensor([roles], dtype=torch.long))
    rows, tokens = [], [0]
    for r in chosen:
        if r == 0:
            break
        rows.append(step_logprobs(W, torch.tensor([tokens], dtype=torch.long), prior)[0])
        tokens.append(int(r))
    return rows


def reconstruct_tiles(final_rank: Sequence[int], tiles: np.ndarray, verb_list: np.ndarray):
    """eval_coco.py:217-237 literally, on one caption's materialised slot tiles (fixed_len, R, F) and verb list (fixed_len, 1):
    permutation matrix, matrix product, rows with a zero sum dropped, tail filled with the last kept row; the verb list goes
    through the same matrix, -1 on the rows nothing was placed in."""
    fixed_len = tiles.shape[0]
    perm_matrix = np.zeros((fixed_len, fixed_len))
    for j, rk in enumerate(final_rank):
        if j < fixed_len:
            perm_matrix[j, int(rk)] = 1
    recons = np.dot(perm_matrix, tiles.reshape(fixed_len, -1)).reshape(tiles.shape)
    recons = recons[np.sum(recons, (1, 2)) != 0]
    out = np.zeros(tiles.shape)
    last = recons.shape[0] - 1
    out[:recons.shape[0]] = recons
    out[last + 1:] = recons[last:last + 1]
    perm_mask = (np.sum(perm_matrix, -1) == 0).astype(int)
    verbs = -1 * perm_mask[:, np.newaxis] + np.dot(perm_matrix, verb_list.reshape(fixed_len, 1))
    return out, verbs[:, 0]
