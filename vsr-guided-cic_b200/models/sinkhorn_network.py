"""R-level SSP network (SinkhornNet), B200-native drop-in.

Keeps the reference's class surface (models/sinkhorn_network.py:5-51): constructor `SinkhornNet(N, n_iters, tau)`, the
five `nn.Linear` layers under the same names (same state_dict, so `load_state_dict(torch.load(...))` of the published
R-level SSP checkpoint works), `forward(seq)` -> (B, N, N) doubly-stochastic matrices.  The arithmetic runs in libvsrdec's
`k_sinkhorn` kernel (one CTA per problem: MLP, Sinkhorn iterations and — through `assign()` — the Hungarian algorithm
that the reference runs on the host with munkres after a `.cpu()` per repeated role, coco_scripts/eval_coco.py:184-189).
CUDA tensors only: this path has no CPU fallback."""
import ctypes

import torch
from torch import nn

_PARAMS = ("W1_txt.weight", "W1_txt.bias", "W1_vis.weight", "W1_vis.bias", "W2_vis.weight", "W2_vis.bias",
           "W_fc_pos.weight", "W_fc_pos.bias", "W_fc.weight", "W_fc.bias")


class SinkhornNet(nn.Module):
    def __init__(self, N, n_iters, tau):
        super().__init__()
        self.N = N
        self.n_iters = n_iters
        self.tau = tau
        self.W1_txt = nn.Linear(300, 128)
        self.W1_vis = nn.Linear(2048, 512)
        self.W2_vis = nn.Linear(512, 128)
        self.W_fc_pos = nn.Linear(260, 256)
        self.W_fc = nn.Linear(256, N)
        self._handle = None
        self._key = None
        self.init_weights()

    def init_weights(self):
        """(reference sinkhorn_network.py:18-28) Xavier-normal weights, zero biases, in the same order."""
        for lin in (self.W1_txt, self.W1_vis, self.W2_vis, self.W_fc_pos, self.W_fc):
            nn.init.xavier_normal_(lin.weight)
            nn.init.constant_(lin.bias, 0)

    # ------------------------------------------------------------------ device handle
    def _engine(self):
        from vsrdec import _lib
        lib = _lib.load_library()
        sd = dict(self.named_parameters())
        ws = [sd[n].detach().contiguous() for n in _PARAMS]
        if not ws[0].is_cuda:
            raise _lib.VsrError("SinkhornNet: parameters are on the CPU; this path runs only on a CUDA device "
                                "(no CPU fallback) — call .to('cuda') first")
        key = tuple((w.data_ptr(), w._version, w.device) for w in ws)
        arr = (_lib.c_vp * 10)(*[w.data_ptr() for w in ws])
        with torch.cuda.device(ws[0].device):
            if self._handle is None or self._key is None or self._key[0][2] != ws[0].device:
                self.close()
                h = _lib.c_vp()
                torch.cuda.current_stream(ws[0].device).synchronize()
                _lib.check(lib, lib.vsr_ssp_create(arr, int(self.N), int(self.n_iters), float(self.tau), ctypes.byref(h)))
                self._handle = h
            elif key != self._key:
                _lib.check(lib, lib.vsr_ssp_load_weights(self._handle, arr, torch.cuda.current_stream(ws[0].device).cuda_stream))
        self._key = key
        return lib

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            from vsrdec import _lib
            _lib.load_library().vsr_ssp_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ forward / assignment
    def _run(self, seq, want_assign):
        from vsrdec import _lib
        if not isinstance(seq, torch.Tensor) or not seq.is_cuda:
            raise _lib.VsrError("SinkhornNet: seq must be a CUDA tensor (no CPU fallback on this path)")
        if seq.dim() != 3 or seq.size(1) != self.N or seq.size(2) != 2352:
            raise _lib.VsrError(f"SinkhornNet: seq must be (B, {self.N}, 2352), got {tuple(seq.shape)}")
        lib = self._engine()
        x = seq.float().contiguous()
        B = x.size(0)
        matrix = torch.empty((B, self.N, self.N), device=x.device, dtype=torch.float32)
        assign = torch.empty((B, self.N), device=x.device, dtype=torch.int32) if want_assign else None
        with torch.cuda.device(x.device):
            _lib.check(lib, lib.vsr_ssp_forward(self._handle, x.data_ptr(), B, matrix.data_ptr(),
                                                assign.data_ptr() if want_assign else None,
                                                torch.cuda.current_stream(x.device).cuda_stream))
        return matrix, assign

    def forward(self, seq):
        """(reference sinkhorn_network.py:39-51) seq (B, N, 2352) -> (B, N, N)."""
        return self._run(seq, False)[0]

    def assign(self, seq):
        """Extension: (matrix (B,N,N), assign (B,N) int32) where assign[b, r] is the column that the maximum-profit
        assignment of matrix[b]^T gives row r — the pairs (r, col) that `munkres.Munkres().compute(make_cost_matrix(mx))`
        returns in eval_coco.py:187-189, computed on the device for the whole batch of problems at once."""
        return self._run(seq, True)
