"""S-level SSP network (S_SSP: the semantic-role sorter of the eval pre-step), B200-native drop-in.

Keeps the reference's class surface (models/sort_model.py:13-50): constructor `S_SSP(pos_enc=False, add_fc=True,
dataset='coco')`, the same module tree and therefore the same 145-key state_dict (so
`load_state_dict(torch.load('saved_model/coco_s_ssp/model-tr.pth'))`, coco_scripts/eval_coco.py:95-97, works), the same
seeded initialisation (constructor order + Xavier-uniform over `parameters()`, sort_model.py:16-17, 52-55 — checked bit for
bit against the reference by tests/test_oracle_sort.py), and

    generate(this_verb, det_seqs_sr, mode='not-normal') -> (pred (1, 10) long, seqLogprobs (1, 10), None)

as eval_coco.py:174 calls it.  The arithmetic — 3-layer encoder over the verb + role embeddings, 3-layer decoder with a
key/value cache, the greedy choice among the roles still to be placed — runs in libvsrdec (`vsr_sort_generate`, csrc/sort.cu)
for a whole batch of (verb, role set) problems per call: `generate_batch(verbs (P,), roles (P, 10))`.  CUDA tensors only: no
CPU fallback.  Training (`forward`, the label-smoothed loss), the unconstrained `mode='normal'` and the dead beam-search branch
(`beam_size` is fixed to 1, sort_model.py:27) are outside this path and raise."""
import ctypes
import math

import torch
from torch import nn


class _PositionalTable(nn.Module):
    """(transformer_modules.py:278-296) only its `pe` buffer matters here: it is part of the checkpoint's state_dict."""

    def __init__(self, size, max_len=5000):
        super().__init__()
        pe = torch.zeros(max_len, size)
        position = torch.arange(0, max_len).unsqueeze(1).float()
        div_term = torch.exp((torch.arange(0, size, 2).float() * -(math.log(10000.0) / size)).float())
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class _Embedding(nn.Embedding):
    def __init__(self, num_embeddings, embedding_dim):
        super().__init__(num_embeddings, embedding_dim)
        self.pos_layer = _PositionalTable(embedding_dim)


class _Attention(nn.Module):
    def __init__(self, size):
        super().__init__()
        self.linear_Q = nn.Linear(size, size)
        self.linear_K = nn.Linear(size, size)
        self.linear_V = nn.Linear(size, size)
        self.linear_O = nn.Linear(size, size)


class _FeedForward(nn.Module):
    def __init__(self, size, hidden):
        super().__init__()
        self.w_1 = nn.Linear(size, hidden)
        self.w_2 = nn.Linear(hidden, size)


class _EncoderLayer(nn.Module):
    def __init__(self, size, ff_size):
        super().__init__()
        self.attention = _Attention(size)
        self.ff_layer = _FeedForward(size, ff_size)
        self.layer_norm1 = nn.LayerNorm(size)
        self.layer_norm2 = nn.LayerNorm(size)


class _DecoderLayer(nn.Module):
    def __init__(self, size, ff_size):
        super().__init__()
        self.attention = _Attention(size)
        self.cross_attention = _Attention(size)      # in the checkpoint, never used by the reference's forward (sort_modules.py:88)
        self.ff_layer = _FeedForward(size, ff_size)
        self.layer_norm1 = nn.LayerNorm(size)
        self.layer_norm2 = nn.LayerNorm(size)
        self.layer_norm3 = nn.LayerNorm(size)


class _Encoder(nn.Module):
    def __init__(self, sr_embed_layer, v_embed_layer, size, n_layers, add_fc):
        super().__init__()
        self.sr_embed_layer = sr_embed_layer
        self.v_embed_layer = v_embed_layer
        self.layer_norm = nn.LayerNorm(size)
        self.encoder_layers = nn.ModuleList()
        if add_fc:
            self.fc_feat = nn.Linear(size, size)
        for _ in range(n_layers):
            self.encoder_layers.append(_EncoderLayer(size, size * 4))


class _Decoder(nn.Module):
    def __init__(self, embed_layer, size, n_layers):
        super().__init__()
        self.embed_layer = embed_layer
        self.layer_norm = nn.LayerNorm(size)
        self.encoder_layers = nn.ModuleList()
        for _ in range(n_layers):
            self.encoder_layers.append(_DecoderLayer(size, size * 4))


class _LabelSmoothing(nn.Module):
    def __init__(self, label_smoothing, n):
        super().__init__()
        self.register_buffer("one_hot", torch.full((n,), label_smoothing / (n - 2)).unsqueeze(0))


def _att(prefix):
    return [f"{prefix}.linear_{m}.{p}" for m in "QKVO" for p in ("weight", "bias")]


def _ff_ln(prefix, n_ln):
    names = [f"{prefix}.ff_layer.{m}.{p}" for m in ("w_1", "w_2") for p in ("weight", "bias")]
    return names + [f"{prefix}.layer_norm{i}.{p}" for i in range(1, n_ln + 1) for p in ("weight", "bias")]


class S_SSP(nn.Module):
    N_ROLES = 26

    def __init__(self, pos_enc=False, add_fc=True, dataset='coco'):
        super().__init__()
        torch.manual_seed(1234)
        if torch.cuda.is_available():
            torch.cuda.manual_seed(1234)
        self._verb_size = 2662 if dataset == 'coco' else 2926
        self.encoder_layers = 3
        self.decoder_layers = 3
        self.max_len = 10
        self.beam_size = 1
        self.hidden_size = 512
        self.embed_size = 512
        self.pos_enc, self.add_fc = bool(pos_enc), bool(add_fc)
        self.sr_embed_layer = _Embedding(self.N_ROLES, self.embed_size)
        self.v_embed_layer = _Embedding(self._verb_size + 1, self.embed_size)
        self.encoder = _Encoder(self.sr_embed_layer, self.v_embed_layer, self.hidden_size, self.encoder_layers, add_fc)
        self.decoder = _Decoder(self.sr_embed_layer, self.hidden_size, self.decoder_layers)
        self.expander_nn = nn.Linear(self.hidden_size, self.N_ROLES)
        self.label_smooth = _LabelSmoothing(0.1, self.N_ROLES)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self._handle = None
        self._key = None

    # ------------------------------------------------------------------ device handle
    def _weight_names(self):
        """Order of the weight pointers in the C ABI (include/vsrdec.h, vsr_sort_create)."""
        names = ["sr_embed_layer.weight", "v_embed_layer.weight",
                 "encoder.fc_feat.weight" if self.add_fc else None, "encoder.fc_feat.bias" if self.add_fc else None,
                 "encoder.layer_norm.weight", "encoder.layer_norm.bias"]
        for l in range(self.encoder_layers):
            p = f"encoder.encoder_layers.{l}"
            names += _att(p + ".attention") + _ff_ln(p, 2)
        names += ["decoder.layer_norm.weight", "decoder.layer_norm.bias"]
        for l in range(self.decoder_layers):
            p = f"decoder.encoder_layers.{l}"
            names += _att(p + ".attention") + _ff_ln(p, 3)
        return names + ["expander_nn.weight", "expander_nn.bias"]

    def _engine(self):
        from vsrdec import _lib
        if self.pos_enc:
            raise _lib.VsrError("S_SSP: pos_enc=True is not on this path (the published eval builds S_SSP() without it)")
        lib = _lib.load_library()
        sd = dict(self.named_parameters())
        names = self._weight_names()
        ws = [None if n is None else sd[n].detach().contiguous() for n in names]
        if not ws[0].is_cuda:
            raise _lib.VsrError("S_SSP: parameters are on the CPU; this path runs only on a CUDA device (no CPU fallback) "
                                "— call .cuda() first")
        dev = ws[0].device
        key = tuple((0, 0, dev) if w is None else (w.data_ptr(), w._version, w.device) for w in ws)
        arr = (_lib.c_vp * len(ws))(*[None if w is None else w.data_ptr() for w in ws])
        with torch.cuda.device(dev):
            if self._handle is None or self._key is None or self._key[0][2] != dev:
                self.close()
                dims = _lib.VsrSortDims(self.N_ROLES, self._verb_size + 1, self.hidden_size, self.hidden_size * 4, 8,
                                        self.encoder_layers, self.max_len, int(self.add_fc))
                h = _lib.c_vp()
                torch.cuda.current_stream(dev).synchronize()
                _lib.check(lib, lib.vsr_sort_create(ctypes.byref(dims), arr, len(ws), ctypes.byref(h)))
                self._handle = h
            elif key != self._key:
                _lib.check(lib, lib.vsr_sort_load_weights(self._handle, arr, len(ws), torch.cuda.current_stream(dev).cuda_stream))
        self._key = key
        return lib

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            from vsrdec import _lib
            _lib.load_library().vsr_sort_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ generation
    def generate_batch(self, verbs, roles, n_steps=None, trace=False, counts=None):
        """P independent problems in one call.  verbs (P,) and roles (P, max_len): CUDA integer tensors; a problem's roles are
        its distinct non-zero role ids, zero-padded.  Returns (pred (P, max_len) long, seqLogprobs (P, max_len) float) — the
        role ids in generated order and the log-prob of each choice, zero after the last role — and, with trace=True, the
        (P, n_steps, 26) log-prob rows of every step.  n_steps: number of decoder steps to run (default max_len; the largest
        role count of the batch is enough).  counts: host sequence of the P role counts, if the caller knows them: the problems
        are then decoded in order of falling count and every decoder step runs only on the problems that still have a role to
        place (results are returned in the caller's order)."""
        from vsrdec import _lib
        if not (isinstance(verbs, torch.Tensor) and isinstance(roles, torch.Tensor) and verbs.is_cuda and roles.is_cuda):
            raise _lib.VsrError("S_SSP: verbs and roles must be CUDA tensors (no CPU fallback on this path)")
        if roles.dim() != 2 or roles.size(1) != self.max_len or verbs.numel() != roles.size(0):
            raise _lib.VsrError(f"S_SSP: expected verbs (P,) and roles (P, {self.max_len}), got {tuple(verbs.shape)} and "
                                f"{tuple(roles.shape)}")
        lib = self._engine()
        P = roles.size(0)
        verbs = (verbs.reshape(-1) % 10000).long().contiguous()       # sort_model.py:108
        roles = roles.long().contiguous()
        order = active = None
        if counts is not None and P:
            counts = [int(x) for x in counts]
            if len(counts) != P:
                raise _lib.VsrError(f"S_SSP: counts has {len(counts)} entries for {P} problems")
            idx = sorted(range(P), key=lambda i: -counts[i])
            order = torch.tensor(idx, dtype=torch.long, device=roles.device)
            verbs, roles = verbs[order].contiguous(), roles[order].contiguous()
            n_steps = min(self.max_len, max(counts)) if n_steps is None else int(n_steps)
            active = (ctypes.c_int32 * max(n_steps, 1))(*[sum(1 for c in counts if c > t) for t in range(n_steps)])
        n_steps = self.max_len if n_steps is None else int(n_steps)
        pred = torch.zeros((P, self.max_len), dtype=torch.long, device=roles.device)
        logp = torch.zeros((P, self.max_len), dtype=torch.float32, device=roles.device)
        rows = torch.zeros((P, n_steps, self.N_ROLES), dtype=torch.float32, device=roles.device) if trace else None
        if P:
            with torch.cuda.device(roles.device):
                st = torch.cuda.current_stream(roles.device).cuda_stream
                _lib.check(lib, lib.vsr_sort_generate(self._handle, verbs.data_ptr(), roles.data_ptr(), P, n_steps, active, pred.data_ptr(),
                                                      logp.data_ptr(), None if rows is None else rows.data_ptr(), st))
        if order is not None:
            inv = torch.empty_like(order)
            inv[order] = torch.arange(P, device=order.device)
            pred, logp = pred[inv], logp[inv]
            rows = rows[inv] if rows is not None else None
        return (pred, logp, rows) if trace else (pred, logp)

    def generate(self, this_verb, det_seqs_sr, mode='normal'):
        """(reference sort_model.py:105-183)  this_verb (1,) / (1, 1), det_seqs_sr (1, max_len)."""
        from vsrdec import _lib
        if mode == 'normal':
            raise _lib.VsrError("S_SSP.generate: only mode='not-normal' (the constrained order the eval uses, eval_coco.py:174) "
                                "is on this path")
        pred, logp = self.generate_batch(this_verb.reshape(-1)[:1], det_seqs_sr.reshape(1, -1))
        return pred, logp, None

    def forward(self, this_verb, det_seqs_sr, gt_seqs_sr):
        from vsrdec import _lib
        raise _lib.VsrError("S_SSP.forward (the training loss) is outside the accelerated path")
