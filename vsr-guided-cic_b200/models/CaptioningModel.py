"""Decode drivers of the captioning model, B200-native.

Same class surface as the reference's models/CaptioningModel.py:8-294 (`forward`, `test`,
`sample_rl`, `beam_search`, `beam_search_v`, `_select_beam[_i]`), but each driver hands the
WHOLE loop to libvsrdec (vsr_forward_teacher / vsr_greedy / vsr_sample / vsr_beam_search): the per-step
Python loop, the per-step statics gather, the full candidate sort and every host sync of the
reference are gone.  Sub-classes provide `_engine_for(statics, seqs, verbs)`.
"""
import torch
from torch import nn


class CaptioningModel(nn.Module):
    def __init__(self, seq_len):
        self.seq_len = seq_len
        super().__init__()

    # ---- interface of the reference base class (CaptioningModel.py:13-20)
    def init_weights(self):
        raise NotImplementedError

    def init_state(self, b_s, device):
        raise NotImplementedError

    def step(self, t, state, prev_outputs, images, seqs, *args, mode='teacher_forcing'):
        raise NotImplementedError

    def _engine_for(self, statics, seqs=None):
        raise NotImplementedError

    # ---- drivers
    def forward(self, statics, seqs, *args):
        """Teacher-forced unroll (reference CaptioningModel.py:22-36): returns
        (out (B,T,V), gate (B,T,2)) log-probabilities."""
        eng = self._engine_for(statics, seqs)
        return eng.forward_teacher(seqs[0])

    def test(self, statics, *args):
        """Greedy decode (reference CaptioningModel.py:38-52): (words (b,T), gates (b,T))."""
        eng = self._engine_for(statics)
        return eng.greedy()

    def sample_rl(self, statics, *args, seed=None):
        """Multinomial sampling with log-probs (reference CaptioningModel.py:54-76): at every step both heads are
        sampled from the step's distributions and fed back; returns ((words, gates), (lp_words, lp_gates)), each (b,T).
        The whole loop runs on the device (vsr_sample: Gumbel-max over the vocabulary row with a Philox stream).  The
        stream's seed is drawn from torch's CPU generator unless given, so torch.manual_seed() makes it reproducible
        (the draws differ from torch.distributions' — another generator — but follow the same distributions)."""
        eng = self._engine_for(statics)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        return eng.sample(seed)

    def _select_beam(self, input, selected_beam, cur_beam_size, beam_size, b_s, reduced=True):
        """Beam-axis gather over (nested) tensors (reference CaptioningModel.py:78-94).  Kept for
        API compatibility; the device path never copies statics per beam."""
        if not isinstance(input, (list, tuple)):
            return self._select_beam_i(input, selected_beam, cur_beam_size, beam_size, b_s, reduced=reduced)
        picked = []
        for s in input:
            if isinstance(s, (list, tuple)):
                picked.append(tuple(self._select_beam_i(ss, selected_beam, cur_beam_size, beam_size, b_s,
                                                        reduced=reduced) for ss in s))
            else:
                picked.append(self._select_beam_i(s, selected_beam, cur_beam_size, beam_size, b_s,
                                                  reduced=reduced))
        return picked

    def _select_beam_i(self, input, selected_beam, cur_beam_size, beam_size, b_s, reduced=True):
        """(reference CaptioningModel.py:96-114)"""
        tail = tuple(input.shape[1:] if reduced else input.shape[2:])
        src = input.reshape((b_s, cur_beam_size) + tail)
        rows = torch.arange(b_s, device=input.device).unsqueeze(1)
        out = src[rows, selected_beam.long().reshape(b_s, beam_size)]
        return out.reshape((b_s * beam_size,) + tail) if reduced else out

    def _beam(self, statics, eos_idxs, beam_size, out_size, use_verbs, gt):
        eng = self._engine_for(statics)
        (words, gates), (lpw, lpg), _ = eng.beam_search(beam_size, out_size, eos_idxs,
                                                        use_verbs=use_verbs, gt=gt)
        outputs, log_probs = [words, gates], [lpw, lpg]
        if out_size == 1:
            outputs = [o.squeeze(1) for o in outputs]
            log_probs = [lp.squeeze(1) for lp in log_probs]
        return outputs, log_probs

    def beam_search_v_indexed(self, statics, eos_idxs, beam_size, out_size=1, *args, gt=False):
        """Extension: beam_search_v with the slots given as INDICES into the detections instead
        of materialised feature tiles: statics = (detections (b,D,F), slot_index (b,L,R) int, verbs (b,L) or None);
        slot_index >= 0 picks a detection row, -2 the mean of the image's valid detections, -1 is padding.
        Same return structure as beam_search_v; 10x fewer input bytes."""
        eng = self._engine()
        verbs = statics[2] if len(statics) > 2 else None
        eng.prologue_indexed(statics[0], statics[1], verbs)
        self._prologue_key = None
        (words, gates), (lpw, lpg), _ = eng.beam_search(beam_size, out_size, eos_idxs, use_verbs=verbs is not None, gt=gt)
        outputs, log_probs = [words, gates], [lpw, lpg]
        if out_size == 1:
            outputs = [o.squeeze(1) for o in outputs]
            log_probs = [lp.squeeze(1) for lp in log_probs]
        return outputs, log_probs

    def beam_search(self, statics, eos_idxs, beam_size, out_size=1, *args):
        """Joint (word, gate) beam search (reference CaptioningModel.py:116-195).  beam_size <= 8, out_size <= beam_size,
        <= 64 regions per slot (see the limits in models/controllable_captioning.py)."""
        return self._beam(statics[:2], eos_idxs, beam_size, out_size, False, False)

    def beam_search_v(self, statics, eos_idxs, beam_size, out_size=1, *args, gt=False):
        """Beam search with verb forcing (reference CaptioningModel.py:197-294).  beam_size <= 8, out_size <= beam_size,
        <= 64 regions per slot (see the limits in models/controllable_captioning.py)."""
        return self._beam(statics, eos_idxs, beam_size, out_size, True, gt)
