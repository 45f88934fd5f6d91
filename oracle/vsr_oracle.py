"""CPU oracle for the role-shift captioning decoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vsr-guided-cic_b200/`` may import this
module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker (and
as the timed CPU port), never as the product path.

What it is: a functional restatement (explicit weight dict, plain torch CPU
ops, no nn.Module) of
  * ``ControllableCaptioningModel.step`` / ``step_v``
    (reference ``models/controllable_captioning.py:117-190`` / ``:192-297``),
  * ``CaptioningModel.forward`` / ``beam_search`` / ``beam_search_v`` /
    ``test`` (reference ``models/CaptioningModel.py:22-36, 116-195, 197-294, 38-52``),
  * the constructor's parameter set and ``init_weights``
    (``controllable_captioning.py:11-107``).
All arithmetic of the reference lives in un-vendored PyTorch (1.5.1 pinned in
``vsr.yml:140``; torch 2.11 is the executable spec here), so the oracle calls
the same torch CPU primitives in the same association order; everything else
(control flow, bookkeeping, history) is written from the behaviour.

Parity pin: the reference ships no tests/golden vectors of its own
(SURVEY.md §4, §8c).  The pin is ``tests/golden/*.pt``: outputs of the
UNMODIFIED reference imported from /root/reference in the build container by
``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks this
oracle against them bit-for-bit (tokens) / to <=1e-6 (log-probs), and
``tests/test_oracle_vs_reference.py`` re-checks live whenever /root/reference
is present.
"""
from __future__ import annotations

import collections
import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- dims / params

@dataclass(frozen=True)
class Dims:
    """Constructor arguments of the reference model (controllable_captioning.py:11-12)."""
    seq_len: int = 20
    vocab_size: int = 10000
    bos_idx: int = 2
    det_feat_size: int = 2048
    input_encoding_size: int = 1000
    rnn_size: int = 1000
    att_size: int = 512
    h2_first_lstm: bool = True
    img_second_lstm: bool = False

    @property
    def in1(self) -> int:  # width of input_1 (controllable_captioning.py:36-39, 228-230)
        w = self.det_feat_size + self.input_encoding_size
        return w + self.rnn_size if self.h2_first_lstm else w

    @property
    def in2(self) -> int:  # width of input_2 (controllable_captioning.py:54-57, 254-257)
        w = self.rnn_size + self.det_feat_size
        return w + self.det_feat_size if self.img_second_lstm else w

    def asdict(self):
        return asdict(self)


def param_shapes(d: Dims) -> "collections.OrderedDict[str, Tuple[int, ...]]":
    """The 28 state_dict entries in registration order (controllable_captioning.py:23-68)."""
    H, E, Fd, A, V = d.rnn_size, d.input_encoding_size, d.det_feat_size, d.att_size, d.vocab_size
    s = collections.OrderedDict()
    s["embed.weight"] = (V, E)
    s["W1_is.weight"] = (H, d.in1); s["W1_is.bias"] = (H,)
    s["W1_hs.weight"] = (H, H); s["W1_hs.bias"] = (H,)
    s["att_va.weight"] = (A, Fd)
    s["att_ha.weight"] = (A, H)
    s["att_a.weight"] = (1, A)
    s["att_sa.weight"] = (A, H)
    s["att_s.weight"] = (1, A)
    s["lstm_cell_1.weight_ih"] = (4 * H, d.in1); s["lstm_cell_1.weight_hh"] = (4 * H, H)
    s["lstm_cell_1.bias_ih"] = (4 * H,); s["lstm_cell_1.bias_hh"] = (4 * H,)
    s["lstm_cell_2.weight_ih"] = (4 * H, d.in2); s["lstm_cell_2.weight_hh"] = (4 * H, H)
    s["lstm_cell_2.bias_ih"] = (4 * H,); s["lstm_cell_2.bias_hh"] = (4 * H,)
    s["out_fc.weight"] = (V, H); s["out_fc.bias"] = (V,)
    s["s_fc.weight"] = (Fd, H); s["s_fc.bias"] = (Fd,)
    s["W1_ig.weight"] = (H, d.in1); s["W1_ig.bias"] = (H,)
    s["W1_hg.weight"] = (H, H); s["W1_hg.bias"] = (H,)
    s["att_ga.weight"] = (A, H)
    s["att_g.weight"] = (1, A)
    return s


PARAM_NAMES = tuple(param_shapes(Dims()).keys())


def init_weights(d: Dims, seed: Optional[int] = 1234) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's initialisers, consuming the torch
    global RNG the way the reference constructor does so that, for the same seed
    and torch build, the tensors are bit-identical to
    ``ControllableCaptioningModel(...).state_dict()``.

    The constructor first default-initialises each layer in declaration order
    (controllable_captioning.py:23-68) and then ``init_weights`` re-draws every
    matrix (``:72-107``): Xavier-normal for all Linear/Embedding matrices and LSTM
    ``weight_ih``, orthogonal for LSTM ``weight_hh``, zeros for all biases.
    """
    if seed is not None:
        torch.manual_seed(seed)
    nn = torch.nn
    H, E, Fd, A, V = d.rnn_size, d.input_encoding_size, d.det_feat_size, d.att_size, d.vocab_size
    # declaration order == RNG consumption order of the default initialisers
    decl = collections.OrderedDict()
    decl["embed"] = nn.Embedding(V, E)
    decl["W1_is"] = nn.Linear(d.in1, H)
    decl["W1_hs"] = nn.Linear(H, H)
    decl["att_va"] = nn.Linear(Fd, A, bias=False)
    decl["att_ha"] = nn.Linear(H, A, bias=False)
    decl["att_a"] = nn.Linear(A, 1, bias=False)
    decl["att_sa"] = nn.Linear(H, A, bias=False)
    decl["att_s"] = nn.Linear(A, 1, bias=False)
    decl["lstm_cell_1"] = nn.LSTMCell(d.in1, H)
    decl["lstm_cell_2"] = nn.LSTMCell(d.in2, H)
    decl["out_fc"] = nn.Linear(H, V)
    decl["s_fc"] = nn.Linear(H, Fd)
    decl["W1_ig"] = nn.Linear(d.in1, H)
    decl["W1_hg"] = nn.Linear(H, H)
    decl["att_ga"] = nn.Linear(H, A, bias=False)
    decl["att_g"] = nn.Linear(A, 1, bias=False)
    W = {}
    for mod_name, mod in decl.items():
        for p_name, p in mod.named_parameters():
            W[f"{mod_name}.{p_name}"] = p.data
    # re-initialisation order of init_weights (controllable_captioning.py:72-107)
    redo = ["embed.weight", "out_fc.weight", "s_fc.weight", "W1_is.weight", "W1_hs.weight",
            "att_va.weight", "att_ha.weight", "att_a.weight", "att_sa.weight", "att_s.weight",
            "lstm_cell_1.weight_ih", "lstm_cell_1.weight_hh",
            "lstm_cell_2.weight_ih", "lstm_cell_2.weight_hh",
            "W1_ig.weight", "W1_hg.weight", "att_ga.weight", "att_g.weight"]
    for name in redo:
        if name.endswith("weight_hh"):
            nn.init.orthogonal_(W[name])
        else:
            nn.init.xavier_normal_(W[name])
    for name in W:
        if "bias" in name:
            W[name].zero_()
    out = collections.OrderedDict((k, W[k].detach().clone()) for k in param_shapes(d))
    for k, shp in param_shapes(d).items():
        assert tuple(out[k].shape) == shp, (k, out[k].shape, shp)
    return out


def cast_weights(W: Dict[str, torch.Tensor], dtype) -> Dict[str, torch.Tensor]:
    return collections.OrderedDict((k, v.to(dtype)) for k, v in W.items())


# --------------------------------------------------------------------------- one decoder step

def init_state(d: Dims, n: int, dtype=torch.float32):
    """Zero LSTM states and slot pointer (controllable_captioning.py:109-115)."""
    z = lambda: torch.zeros((n, d.rnn_size), dtype=dtype)
    return (z(), z()), (z(), z()), torch.zeros((n,), dtype=torch.long)


def _lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.LSTMCell forward, gate order i,f,g,o (the reference uses nn.LSTMCell:
    controllable_captioning.py:50-57, 233, 258)."""
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, 1)
    i = torch.sigmoid(i); f = torch.sigmoid(f); g = torch.tanh(g); o = torch.sigmoid(o)
    c_new = f * c + i * g
    h_new = o * torch.tanh(c_new)
    return h_new, c_new


def verb_forced_index(out_row: torch.Tensor, verb: int, gt: bool,
                      verb_table: Optional[Dict[str, List[int]]]) -> int:
    """Vocabulary index that a verb slot forces (controllable_captioning.py:278-292)."""
    if gt:
        return int(verb)
    cands = (verb_table or {}).get(str(int(verb)), [])
    if len(cands) == 0:
        return 0
    best, best_idx = -1e6, -1
    for idx in cands:  # strict '>' : first maximum wins; -1 (last vocab entry) if none beats -1e6
        v = float(out_row[idx])
        if v > best:
            best, best_idx = v, idx
    return best_idx


def decoder_step(W, d: Dims, t: int, state, prev_outputs, statics, seqs, mode: str,
                 use_verbs: bool = False, gt: bool = False, verb_table=None):
    """One step of the role-shift decoder for N independent rows.

    ``use_verbs=False`` follows ``step`` (controllable_captioning.py:117-190);
    ``use_verbs=True`` follows ``step_v`` (``:192-297``), i.e. adds verb forcing.
    Returns ``(out (N,V), gate (N,2)), ((h1,c1),(h2,c2),ptr)``.
    """
    if mode not in ("teacher_forcing", "feedback"):
        raise AssertionError(mode)
    det = statics[0]
    n = det.size(0)
    # image descriptor: sum over all rows / number of rows with non-zero feature sum (:126-128)
    det_mask = (torch.sum(det, -1, keepdim=True) != 0).to(det.dtype)
    img = torch.sum(det, 1) / torch.sum(det_mask, 1)
    (h1, c1), (h2, c2), ptr = state

    verb_curr = None
    if mode == "teacher_forcing":
        if use_verbs:
            raise NameError("step_v supports feedback mode only (verb_curr undefined, :208-223)")
        word_in = seqs[0][:, t]
        det_curr = seqs[1][:, t]
    else:
        if t == 0:
            word_in = torch.full((n,), d.bos_idx, dtype=torch.long)
        else:
            word_in = prev_outputs[0]
            ptr = torch.clamp(ptr + prev_outputs[1], 0, statics[1].shape[1] - 1)  # :139-140
        rows = torch.arange(n)
        det_curr = statics[1][rows, ptr]                      # slot select (:142)
        if use_verbs:
            verb_curr = statics[2][rows, ptr].long()          # (:221)

    xt = F.embedding(word_in, W["embed.weight"])
    in1 = torch.cat([h2, img, xt], 1) if d.h2_first_lstm else torch.cat([img, xt], 1)

    # sentinel gate uses h1 BEFORE the cell update (:151-152)
    s_gate = torch.sigmoid(F.linear(in1, W["W1_is.weight"], W["W1_is.bias"])
                           + F.linear(h1, W["W1_hs.weight"], W["W1_hs.bias"]))
    h1, c1 = _lstm_cell(in1, h1, c1, W["lstm_cell_1.weight_ih"], W["lstm_cell_1.weight_hh"],
                        W["lstm_cell_1.bias_ih"], W["lstm_cell_1.bias_hh"])
    s_t = s_gate * torch.tanh(c1)
    sentinel = F.linear(s_t, W["s_fc.weight"], W["s_fc.bias"]).unsqueeze(1)

    regions = torch.cat([sentinel, det_curr], 1)                              # (N, R+1, F)
    regions_mask = (torch.sum(regions, -1, keepdim=True) != 0).to(det.dtype)   # (:159)

    ha = F.linear(h1, W["att_ha.weight"])
    det_w = F.linear(torch.tanh(F.linear(det_curr, W["att_va.weight"]) + ha.unsqueeze(1)),
                     W["att_a.weight"])                                        # (N, R, 1)
    sent_w = F.linear(torch.tanh(F.linear(s_t, W["att_sa.weight"]) + ha).unsqueeze(1),
                      W["att_s.weight"])                                       # (N, 1, 1)
    alpha = F.softmax(torch.cat([sent_w, det_w], 1), 1)
    alpha = regions_mask * alpha
    alpha = alpha / torch.sum(alpha, 1, keepdim=True)                          # (:167-169)
    att = torch.sum(regions * alpha, 1)

    in2 = torch.cat([h1, att, img], 1) if d.img_second_lstm else torch.cat([h1, att], 1)
    h2, c2 = _lstm_cell(in2, h2, c2, W["lstm_cell_2.weight_ih"], W["lstm_cell_2.weight_hh"],
                        W["lstm_cell_2.bias_ih"], W["lstm_cell_2.bias_hh"])
    out = F.log_softmax(F.linear(h2, W["out_fc.weight"], W["out_fc.bias"]), dim=-1)

    # shift gate uses h1 AFTER the update (:181-188)
    g_gate = torch.sigmoid(F.linear(in1, W["W1_ig.weight"], W["W1_ig.bias"])
                           + F.linear(h1, W["W1_hg.weight"], W["W1_hg.bias"]))
    g_t = g_gate * torch.tanh(c1)
    stay = F.linear(torch.tanh(F.linear(g_t, W["att_ga.weight"]) + ha).unsqueeze(1), W["att_g.weight"])
    shift = torch.sum(regions_mask[:, 1:] * det_w, 1, keepdim=True)
    gate = F.log_softmax(torch.cat([stay, shift], 1), 1).squeeze(-1)           # (N, 2)

    if use_verbs:
        # verb forcing (:271-295): arithmetic blend with an int mask
        verb_mask = (verb_curr != -1).int().unsqueeze(-1)
        verb_out = torch.ones(out.shape, dtype=out.dtype) * -1e6
        for i in torch.nonzero(verb_curr != -1).flatten().tolist():
            verb_out[i, verb_forced_index(out[i], int(verb_curr[i]), gt, verb_table)] = 0
        change_gate = torch.tensor([-1e3, 0], dtype=out.dtype)
        out = (1 - verb_mask) * out + verb_mask * verb_out
        gate = (1 - verb_mask) * gate + verb_mask * change_gate.unsqueeze(0)

    return (out, gate), ((h1, c1), (h2, c2), ptr)


# --------------------------------------------------------------------------- drivers

def forward_teacher(W, d: Dims, statics, seqs):
    """Teacher-forced unroll (CaptioningModel.py:22-36): out (B,T,V), gate (B,T,2)."""
    b = statics[0].size(0)
    state = init_state(d, b, statics[0].dtype)
    outs, gates = [], []
    prev = None
    for t in range(seqs[0].size(1)):
        prev, state = decoder_step(W, d, t, state, prev, statics, seqs, "teacher_forcing")
        outs.append(prev[0]); gates.append(prev[1])
    return torch.stack(outs, 1), torch.stack(gates, 1)


def greedy_test(W, d: Dims, statics):
    """Greedy decode (CaptioningModel.py:38-52): argmax of both heads each step."""
    b = statics[0].size(0)
    state = init_state(d, b, statics[0].dtype)
    prev = None
    words, gates = [], []
    for t in range(d.seq_len):
        outs, state = decoder_step(W, d, t, state, prev, statics, None, "feedback")
        prev = tuple(torch.max(o, -1)[1] for o in outs)
        words.append(prev[0]); gates.append(prev[1])
    return torch.stack(words, 1), torch.stack(gates, 1)


def _pick_beams(x: torch.Tensor, sel_beam: torch.Tensor, b: int, cur: int, flat: bool):
    """Gather along the beam axis (CaptioningModel.py:78-114): x is (b*cur, ...) when
    ``flat`` else (b, cur, ...); returns (b*k, ...) / (b, k, ...)."""
    k = sel_beam.size(1)
    xv = x.reshape((b, cur) + tuple(x.shape[1:] if flat else x.shape[2:]))
    picked = xv[torch.arange(b).unsqueeze(1), sel_beam]          # (b, k, ...)
    return picked.reshape((b * k,) + tuple(xv.shape[2:])) if flat else picked


@dataclass
class BeamTrace:
    """Per-step record of a beam search, for trajectory replay and tie-band checks."""
    sel_beam: List[torch.Tensor]      # T x (b,k) parent slot
    sel_word: List[torch.Tensor]      # T x (b,k)
    sel_gate: List[torch.Tensor]      # T x (b,k)
    sel_score: List[torch.Tensor]     # T x (b,k) accumulated score after the step
    step_out: List[torch.Tensor]      # T x (N_t,V) word log-probs returned by the step (optional)
    step_gate: List[torch.Tensor]     # T x (N_t,2)
    cand_scores: List[torch.Tensor]   # T x (b, cur*V*2) candidate scores that were sorted (optional)


def beam_search(W, d: Dims, statics, eos_idxs: Sequence[int], beam_size: int, out_size: int = 1,
                use_verbs: bool = False, gt: bool = False, verb_table=None,
                trace: Optional[BeamTrace] = None, keep_step_outputs: bool = False,
                keep_cand_scores: bool = False, forced: Optional[BeamTrace] = None, step_hook=None):
    """Joint (word, gate) beam search.

    ``use_verbs=False`` follows ``beam_search`` (CaptioningModel.py:116-195);
    ``use_verbs=True`` follows ``beam_search_v`` (``:197-294``).  Data movement mirrors the
    reference: the statics of every caption are re-gathered per beam each step
    (``:167-168`` / ``:259-260``) and the candidate set is fully sorted (``:152`` / ``:238``).
    The token history is kept as back-pointers and unrolled at the end, which yields the
    same ``outputs`` as the reference's per-step re-gather of the history (``:170,262``).

    ``forced``: replay the given selections instead of the sorted top-k (trajectory replay for
    per-step parity tests); scores are still taken from this run's own candidates.
    ``step_hook(t, out, gate, flat_candidate_scores)`` is called once per step (streaming checks
    that would not fit in memory as a stored trace).
    """
    k = beam_size
    b = statics[0].size(0)
    dtype = statics[0].dtype
    V2 = None
    state = init_state(d, b, dtype)
    statics = tuple(statics)
    sel_outs = None
    seq_lp = torch.zeros((b, 1, 1, 1), dtype=dtype)
    seq_masks = [torch.ones((b, k), dtype=dtype), torch.ones((b, k), dtype=dtype)]
    parents, words, gates, lp_words, lp_gates = [], [], [], [], []

    for t in range(d.seq_len):
        cur = 1 if t == 0 else k
        (out, gate), state = decoder_step(W, d, t, state, sel_outs, statics, None, "feedback",
                                          use_verbs=use_verbs, gt=gt, verb_table=verb_table)
        V = out.size(-1)
        word_lp = out.view(b, cur, V, 1)
        gate_lp = gate.view(b, cur, 1, 2)
        old = seq_lp
        cand = seq_lp + (word_lp + gate_lp)                    # association order of :139 / :224
        step_word, step_gate = word_lp.reshape(b, cur, V), gate_lp.reshape(b, cur, 2)
        if t > 0:
            live = [(so.view(b, cur) != idx).to(dtype) for idx, so in zip(eos_idxs, sel_outs)]
            seq_masks = [sm * m for sm, m in zip(seq_masks, live)]
            step_word = step_word * seq_masks[0].unsqueeze(-1)
            step_gate = step_gate * seq_masks[1].unsqueeze(-1)
            old = old.expand_as(cand).contiguous()
            old[:, :, 1:] = -999                               # slices the WORD axis (:147 / :232)
            full = torch.clamp(seq_masks[0] + seq_masks[1], 0, 1).view(b, cur, 1, 1)
            cand = full * cand + old * (1 - full)
        flat = cand.view(b, -1)
        if step_hook is not None:
            step_hook(t, out, gate, flat)
        if forced is None:
            sorted_lp, sorted_idx = torch.sort(flat, -1, descending=True)
            top_lp, top_idx = sorted_lp[:, :k], sorted_idx[:, :k]
            sel_beam = top_idx // (V * 2)
            rem = top_idx - sel_beam * (V * 2)
            sel_word = (rem / 2).long()                        # true division + trunc (:164 / :255)
            sel_gate = ((rem - sel_word * 2) / 1).long()
        else:
            sel_beam, sel_word, sel_gate = forced.sel_beam[t], forced.sel_word[t], forced.sel_gate[t]
            top_idx = sel_beam * (V * 2) + sel_word * 2 + sel_gate
            top_lp = torch.gather(flat, 1, top_idx)

        if trace is not None:
            trace.sel_beam.append(sel_beam.clone()); trace.sel_word.append(sel_word.clone())
            trace.sel_gate.append(sel_gate.clone()); trace.sel_score.append(top_lp.clone())
            if keep_step_outputs:
                trace.step_out.append(out.clone()); trace.step_gate.append(gate.clone())
            if keep_cand_scores:
                trace.cand_scores.append(flat.clone())

        # re-gather per-beam state, statics (pure copies, as in the reference) and masks
        (h1, c1), (h2, c2), ptr = state
        state = ((_pick_beams(h1, sel_beam, b, cur, True), _pick_beams(c1, sel_beam, b, cur, True)),
                 (_pick_beams(h2, sel_beam, b, cur, True), _pick_beams(c2, sel_beam, b, cur, True)),
                 _pick_beams(ptr, sel_beam, b, cur, True))
        statics = tuple(_pick_beams(s, sel_beam, b, cur, True) for s in statics)
        seq_masks = [_pick_beams(sm, sel_beam, b, k, False) for sm in seq_masks]
        parents.append(sel_beam); words.append(sel_word); gates.append(sel_gate)
        seq_lp = top_lp.reshape(b, k, 1, 1)
        # per-token log-probs are recorded in slot order and never re-gathered (:175-177 / :271-273)
        lp_words.append(torch.gather(_pick_beams(step_word, sel_beam, b, cur, False), 2,
                                     sel_word.unsqueeze(-1)))
        lp_gates.append(torch.gather(_pick_beams(step_gate, sel_beam, b, cur, False), 2,
                                     sel_gate.unsqueeze(-1)))
        sel_outs = [sel_word.reshape(-1), sel_gate.reshape(-1)]

    # final ordering of beams by score, then unroll back-pointers
    _, order = torch.sort(seq_lp.view(b, k, 1), 1, descending=True)
    order = order.squeeze(-1)                                                    # (b,k)
    T = d.seq_len
    out_words = torch.zeros((b, k, T), dtype=torch.long)
    out_gates = torch.zeros((b, k, T), dtype=torch.long)
    slot = torch.arange(k).unsqueeze(0).expand(b, k).clone()
    for t in range(T - 1, -1, -1):
        out_words[:, :, t] = torch.gather(words[t], 1, slot)
        out_gates[:, :, t] = torch.gather(gates[t], 1, slot)
        slot = torch.gather(parents[t], 1, slot)
    idx = order.unsqueeze(-1).expand(b, k, T)
    out_words = torch.gather(out_words, 1, idx)[:, :out_size]
    out_gates = torch.gather(out_gates, 1, idx)[:, :out_size]
    lpw = torch.gather(torch.cat(lp_words, -1), 1, idx)[:, :out_size]
    lpg = torch.gather(torch.cat(lp_gates, -1), 1, idx)[:, :out_size]
    outs, lps = [out_words, out_gates], [lpw, lpg]
    if out_size == 1:
        outs = [o.squeeze(1) for o in outs]
        lps = [x.squeeze(1) for x in lps]
    return outs, lps


def new_trace() -> BeamTrace:
    return BeamTrace([], [], [], [], [], [], [])


# --------------------------------------------------------------------------- synthetic inputs
# (shared with bench.py's product leg, which may not import oracle/: they live in tools/synth.py)
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from tools.synth import synth_inputs, synth_verb_table, synth_inputs_indexed, materialize_slots  # noqa: E402,F401
