"""Shared helpers for the test-suite (fixture loading, tie-aware comparisons)."""
import os

import torch

from oracle import vsr_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    fx = torch.load(os.path.join(GOLDEN, name), weights_only=False)
    fx["dims_obj"] = O.Dims(**fx["dims"])
    if "det_seqs" in fx and "ctrl_rule" in fx:
        ds = fx["det_seqs"]
        b, L = ds.shape[0], ds.shape[1]
        T = fx["dims"]["seq_len"]
        ctrl = torch.zeros((b, T) + tuple(ds.shape[2:]))
        for t in range(T):
            ctrl[:, t] = ds[:, min(t // 2, L - 1)]
        fx["ctrl"] = ctrl
    return fx


def checksum(t: torch.Tensor) -> float:
    t = t.double().flatten()
    return float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).remainder(97.0)).sum())


def rel_close(a, b, rel=1e-3, abs_floor=1e-4):
    """|a-b| <= rel*|b| + abs_floor element-wise (north-star tolerance for log-probs)."""
    a, b = a.double(), b.double()
    return bool(((a - b).abs() <= rel * b.abs() + abs_floor).all())


def max_rel_err(a, b, abs_floor=1e-4):
    a, b = a.double(), b.double()
    return float(((a - b).abs() / (b.abs() + abs_floor)).max())
