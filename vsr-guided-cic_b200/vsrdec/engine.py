"""DecoderEngine: thin, torch-aware wrapper over the C ABI (one handle per device).

PyTorch is plumbing here (device memory, streams); every hot operation runs in libvsrdec's
sm_100a kernels.  CPU tensors are rejected — there is no fallback path."""
import ctypes
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import VsrError, VsrDims, VsrTrace, c_vp, c_i32, c_i64, c_f

# state_dict registration order of the reference model (controllable_captioning.py:23-68);
# the order of the `weights` array of vsr_create().
PARAM_NAMES = (
    "embed.weight", "W1_is.weight", "W1_is.bias", "W1_hs.weight", "W1_hs.bias", "att_va.weight",
    "att_ha.weight", "att_a.weight", "att_sa.weight", "att_s.weight",
    "lstm_cell_1.weight_ih", "lstm_cell_1.weight_hh", "lstm_cell_1.bias_ih", "lstm_cell_1.bias_hh",
    "lstm_cell_2.weight_ih", "lstm_cell_2.weight_hh", "lstm_cell_2.bias_ih", "lstm_cell_2.bias_hh",
    "out_fc.weight", "out_fc.bias", "s_fc.weight", "s_fc.bias", "W1_ig.weight", "W1_ig.bias",
    "W1_hg.weight", "W1_hg.bias", "att_ga.weight", "att_g.weight")

_VERB_DT = {torch.float64: 0, torch.float32: 1, torch.int64: 2}


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _req_cuda(t: torch.Tensor, name: str, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise VsrError(f"{name} must be a CUDA tensor (no CPU fallback on this path)")
    if dtype is not None and t.dtype != dtype:
        raise VsrError(f"{name} must be {dtype}, got {t.dtype}")
    return t


class DecoderEngine:
    def __init__(self, dims: Dict, weights: Sequence[torch.Tensor]):
        self.lib = _lib.load_library()
        self.dims = dict(dims)
        self.device = weights[0].device
        self._keep = None
        self._weights = None
        self.handle = c_vp()
        d = VsrDims(*(int(self.dims[k]) for k in ("seq_len", "vocab_size", "bos_idx", "det_feat_size",
                                                   "input_encoding_size", "rnn_size", "att_size",
                                                   "h2_first_lstm", "img_second_lstm")))
        arr = self._weight_array(weights)
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()
            _lib.check(self.lib, self.lib.vsr_create(ctypes.byref(d), arr, ctypes.byref(self.handle)))

    def _weight_array(self, weights):
        if len(weights) != len(PARAM_NAMES):
            raise VsrError(f"expected {len(PARAM_NAMES)} weight tensors, got {len(weights)}")
        ws = []
        for name, w in zip(PARAM_NAMES, weights):
            _req_cuda(w, name, torch.float32)
            if w.device != self.device:
                raise VsrError("all parameters must live on one device")
            ws.append(w.detach().contiguous())
        self._weights = ws      # keep the (possibly re-laid-out) tensors alive during packing
        return (c_vp * len(ws))(*[w.data_ptr() for w in ws])

    def load_weights(self, weights):
        arr = self._weight_array(weights)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.vsr_load_weights(self.handle, arr, _stream_ptr(self.device)))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.vsr_destroy(self.handle)
            self.handle = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ verb table
    def set_verb_table(self, table: Optional[Dict[str, List[int]]]):
        """`table` has the JSON shape of verb_2_vob_all: {str(verb_id): [vocab idx, ...]}."""
        items = []
        for k, v in (table or {}).items():
            try:
                key = int(k)
            except ValueError:
                continue   # str(int) keys only can ever match str(verb_curr.item())
            if str(key) != k:
                continue
            items.append((key, [int(x) for x in v]))
        items.sort()
        n = len(items)
        keys = (c_i64 * max(n, 1))(*[k for k, _ in items])
        offs = [0]
        flat = []
        for _, v in items:
            flat.extend(v)
            offs.append(len(flat))
        offsets = (c_i32 * (n + 1))(*offs)
        idx = (c_i32 * max(len(flat), 1))(*flat)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.vsr_set_verb_table(self.handle, keys, offsets, idx, n))

    # ------------------------------------------------------------------ prologue
    def prologue(self, det: torch.Tensor, det_seqs: torch.Tensor, verbs: Optional[torch.Tensor] = None):
        _req_cuda(det, "detections", torch.float32)
        _req_cuda(det_seqs, "det_seqs", torch.float32)
        if det.dim() != 3 or det_seqs.dim() != 4 or det.size(0) != det_seqs.size(0):
            raise VsrError(f"bad static shapes {tuple(det.shape)} / {tuple(det_seqs.shape)}")
        F = self.dims["det_feat_size"]
        if det.size(2) != F or det_seqs.size(3) != F:
            raise VsrError("feature width does not match det_feat_size")
        b, D = det.size(0), det.size(1)
        if b > 1 and det.stride(0) == 0 and det[0].is_contiguous():
            det_k, stride = det[0], 0           # one image expanded over all captions (eval_coco.py:243)
        else:
            det_k = det.contiguous()
            stride = D * F
        ds = det_seqs.contiguous()
        vb, vdt = None, 0
        if verbs is not None:
            _req_cuda(verbs, "verbs")
            if verbs.dtype not in _VERB_DT:
                verbs = verbs.double()
            if tuple(verbs.shape) != (b, ds.size(1)):
                raise VsrError(f"verbs shape {tuple(verbs.shape)} != {(b, ds.size(1))}")
            vb = verbs.contiguous()
            vdt = _VERB_DT[vb.dtype]
        self._keep = (det_k, ds, vb)
        self.b, self.L, self.R = b, ds.size(1), ds.size(2)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.vsr_prologue(
                self.handle, det_k.data_ptr(), stride, D, ds.data_ptr(), b, self.L, self.R,
                vb.data_ptr() if vb is not None else None, vdt, _stream_ptr(self.device)))

    def prologue_indexed(self, det: torch.Tensor, slot_index: torch.Tensor, verbs: Optional[torch.Tensor] = None):
        """Index form of the prologue (vsr_prologue_indexed): slot_index (b, L, R) holds, per region, a
        detection row of the caption's image (>= 0), -2 for the image's mean row, or -1 for padding."""
        _req_cuda(det, "detections", torch.float32)
        _req_cuda(slot_index, "slot_index")
        if det.dim() != 3 or slot_index.dim() != 3 or det.size(0) != slot_index.size(0):
            raise VsrError(f"bad static shapes {tuple(det.shape)} / {tuple(slot_index.shape)}")
        F = self.dims["det_feat_size"]
        if det.size(2) != F:
            raise VsrError("feature width does not match det_feat_size")
        b, D = det.size(0), det.size(1)
        if b > 1 and det.stride(0) == 0 and det[0].is_contiguous():
            det_k, stride = det[0], 0
        else:
            det_k = det.contiguous()
            stride = D * F
        idx = slot_index.to(torch.int32).contiguous()
        vb, vdt = None, 0
        if verbs is not None:
            _req_cuda(verbs, "verbs")
            if verbs.dtype not in _VERB_DT:
                verbs = verbs.double()
            if tuple(verbs.shape) != (b, idx.size(1)):
                raise VsrError(f"verbs shape {tuple(verbs.shape)} != {(b, idx.size(1))}")
            vb = verbs.contiguous()
            vdt = _VERB_DT[vb.dtype]
        self._keep = (det_k, idx, vb)
        self.b, self.L, self.R = b, idx.size(1), idx.size(2)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.vsr_prologue_indexed(
                self.handle, det_k.data_ptr(), stride, D, idx.data_ptr(), b, self.L, self.R,
                vb.data_ptr() if vb is not None else None, vdt, _stream_ptr(self.device)))

    # ------------------------------------------------------------------ single step
    def step(self, h1, c1, h2, c2, slot, word, use_verbs=False, gt=False):
        b, H, V = self.b, self.dims["rnn_size"], self.dims["vocab_size"]
        st = [_req_cuda(x, "state", torch.float32).contiguous() for x in (h1, c1, h2, c2)]
        for x in st:
            if tuple(x.shape) != (b, H):
                raise VsrError(f"state shape {tuple(x.shape)} != {(b, H)}")
        slot = _req_cuda(slot, "slot").long().contiguous()
        word = _req_cuda(word, "word").long().contiguous()
        new = [torch.empty((b, H), device=self.device, dtype=torch.float32) for _ in range(4)]
        out = torch.empty((b, V), device=self.device, dtype=torch.float32)
        gate = torch.empty((b, 2), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.vsr_step(
                self.handle, *[x.data_ptr() for x in st], slot.data_ptr(), word.data_ptr(),
                int(use_verbs), int(gt), *[x.data_ptr() for x in new], out.data_ptr(), gate.data_ptr(),
                _stream_ptr(self.device)))
        return (out, gate), new

    # ------------------------------------------------------------------ decode drivers
    def beam_search(self, beam_size, out_size, eos_idxs, use_verbs=False, gt=False,
                    trace_steps=False, forced=None):
        b, T, V = self.b, self.dims["seq_len"], self.dims["vocab_size"]
        dev = self.device
        words = torch.empty((b, out_size, T), device=dev, dtype=torch.long)
        gates = torch.empty((b, out_size, T), device=dev, dtype=torch.long)
        lpw = torch.empty((b, out_size, T), device=dev, dtype=torch.float32)
        lpg = torch.empty((b, out_size, T), device=dev, dtype=torch.float32)
        eos = (c_i64 * 2)(int(eos_idxs[0]), int(eos_idxs[1]))
        tr, tr_ref, extra = None, None, {}
        if trace_steps or forced is not None:
            tr = VsrTrace()
            if trace_steps:
                extra["step_out"] = torch.zeros((T, b * beam_size, V), device=dev)
                extra["step_gate"] = torch.zeros((T, b * beam_size, 2), device=dev)
                tr.step_out, tr.step_gate = extra["step_out"].data_ptr(), extra["step_gate"].data_ptr()
            if forced is not None:
                fb, fw, fg = [x.to(dev, torch.int32).contiguous() for x in forced]
                for x in (fb, fw, fg):
                    if tuple(x.shape) != (T, b, beam_size):
                        raise VsrError("forced selections must be (T, b, beam)")
                extra["_forced"] = (fb, fw, fg)
                tr.forced_beam, tr.forced_word, tr.forced_gate = fb.data_ptr(), fw.data_ptr(), fg.data_ptr()
            tr_ref = ctypes.byref(tr)
        with torch.cuda.device(dev):
            _lib.check(self.lib, self.lib.vsr_beam_search(
                self.handle, int(beam_size), int(out_size), eos, int(use_verbs), int(gt),
                words.data_ptr(), gates.data_ptr(), lpw.data_ptr(), lpg.data_ptr(), tr_ref,
                _stream_ptr(dev)))
        self._last = (T, b, beam_size)
        return (words, gates), (lpw, lpg), extra

    def history(self):
        T, b, k = self._last
        dev = self.device
        parent, word, gate = [torch.empty((T, b, k), device=dev, dtype=torch.int32) for _ in range(3)]
        score = torch.empty((T, b, k), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(self.lib, self.lib.vsr_get_history(self.handle, parent.data_ptr(), word.data_ptr(),
                                                          gate.data_ptr(), score.data_ptr(), _stream_ptr(dev)))
        return parent, word, gate, score

    def forward_teacher(self, captions: torch.Tensor):
        _req_cuda(captions, "captions", torch.long)
        b, V = self.b, self.dims["vocab_size"]
        if captions.dim() != 2 or captions.size(0) != b:
            raise VsrError("captions must be (b, T)")
        T = captions.size(1)
        cap = captions.contiguous()
        out = torch.empty((b, T, V), device=self.device, dtype=torch.float32)
        gate = torch.empty((b, T, 2), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.vsr_forward_teacher(self.handle, cap.data_ptr(), T, out.data_ptr(),
                                                              gate.data_ptr(), _stream_ptr(self.device)))
        return out, gate

    def greedy(self):
        b, T = self.b, self.dims["seq_len"]
        words = torch.empty((b, T), device=self.device, dtype=torch.long)
        gates = torch.empty((b, T), device=self.device, dtype=torch.long)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.vsr_greedy(self.handle, words.data_ptr(), gates.data_ptr(),
                                                     _stream_ptr(self.device)))
        return words, gates

    def sample(self, seed: int):
        """Multinomial sampling decode (vsr_sample): (words, gates) int64 (b,T), (lp_words, lp_gates) fp32 (b,T)."""
        b, T = self.b, self.dims["seq_len"]
        words = torch.empty((b, T), device=self.device, dtype=torch.long)
        gates = torch.empty((b, T), device=self.device, dtype=torch.long)
        lpw = torch.empty((b, T), device=self.device, dtype=torch.float32)
        lpg = torch.empty((b, T), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.vsr_sample(self.handle, int(seed) & 0xFFFFFFFFFFFFFFFF, words.data_ptr(),
                                                     gates.data_ptr(), lpw.data_ptr(), lpg.data_ptr(),
                                                     _stream_ptr(self.device)))
        return (words, gates), (lpw, lpg)

    # ------------------------------------------------------------------ instrumentation
    def launch_count(self) -> int:
        return int(self.lib.vsr_launch_count(self.handle))

    def gemm_kind(self) -> str:
        return self.lib.vsr_gemm_kind(self.handle).decode()

    def set_profiling(self, on: bool):
        _lib.check(self.lib, self.lib.vsr_set_profiling(self.handle, int(on)))

    def step_times(self):
        """ms of every decoder step of the last profiled beam search (CUDA events between the steps)."""
        cap = 512
        ms = (c_f * cap)()
        with torch.cuda.device(self.device):
            n = self.lib.vsr_get_step_times(self.handle, ms, cap)
        if n < 0:
            _lib.check(self.lib, n)
        return [float(ms[i]) for i in range(n)]

    def phase_times(self):
        cap = 32
        names = (ctypes.c_char_p * cap)()
        ms = (c_f * cap)()
        cnt = (c_i32 * cap)()
        with torch.cuda.device(self.device):
            n = self.lib.vsr_get_phase_times(self.handle, names, ms, cnt, cap)
        if n < 0:
            _lib.check(self.lib, n)
        return [(names[i].decode(), float(ms[i]), int(cnt[i])) for i in range(n)]
