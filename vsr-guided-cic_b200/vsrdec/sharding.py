"""Row (caption) sharding across the GPUs of one host.

Captions are independent units: each rank decodes a contiguous block with its own replica of
the weights, there is no collective on the data path, and the finished captions are gathered
once at the end (all_gather of (b/G, T) tokens + log-probs, ~480 B per caption)."""
from typing import Callable, List, Sequence, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first n_items % world ranks get one extra."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def decode_sharded(decode_fn: Callable[[int, int], Sequence[torch.Tensor]], n_captions: int,
                   rank: int, world: int, group=None) -> List[torch.Tensor]:
    """Run `decode_fn(lo, hi)` on this rank's block and all-gather the per-caption results.

    decode_fn returns tensors whose dim 0 is the caption axis (hi - lo rows).  Every rank
    receives the full (n_captions, ...) tensors in caption order."""
    import torch.distributed as dist
    lo, hi = shard_range(n_captions, rank, world)
    local = [t.contiguous() for t in decode_fn(lo, hi)]
    if world == 1:
        return local
    sizes = [shard_range(n_captions, r, world) for r in range(world)]
    max_rows = max(h - l for l, h in sizes)
    outs = []
    for t in local:
        pad = torch.zeros((max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: hi - lo] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        outs.append(torch.cat([bufs[r][: h - l] for r, (l, h) in enumerate(sizes)], 0))
    return outs
