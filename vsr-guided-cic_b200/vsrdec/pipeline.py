"""Throughput-oriented host loop around the drop-in model's beam search (what `bench.py`'s end-to-end figure runs).

The reference decodes one batch at a time on the default stream and `.cpu()`s the result (eval_coco.py:245-249).
Consecutive batches are independent, so this helper keeps several of them in flight:

* `lanes` decode lanes, each an engine (a model replica with the same weights) on its own CUDA stream: while one
  decode sits in its latency-bound small kernels the other one's GEMMs use the SMs;
* `buffers` device input buffers filled from pinned host memory by a copy stream that runs ahead of the lanes;
* results copied to pinned host memory asynchronously and handed out one step later, in order.

Nothing here changes what a decode computes: every batch goes through `beam_search_v` (or `beam_search_v_indexed`)
exactly as a direct call would."""
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch


class DecodePipeline:
    def __init__(self, models: Sequence, eos_idxs, beam_size: int, out_size: int = 1, gt: bool = False,
                 indexed: bool = False, buffers: Optional[int] = None,
                 post: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        """models: one drop-in model per lane (same weights, same device).  `post` is applied to the words tensor on
        the lane's stream (e.g. the all-gather of a caption-sharded job)."""
        if not models:
            raise ValueError("DecodePipeline needs at least one model")
        self.models = list(models)
        self.device = next(self.models[0].parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("DecodePipeline: the models must live on a CUDA device (no CPU fallback)")
        self.eos_idxs, self.beam_size, self.out_size, self.gt, self.indexed = eos_idxs, beam_size, out_size, gt, indexed
        self.post = post
        self.n_lanes = len(self.models)
        self.n_buf = buffers if buffers is not None else 2 * self.n_lanes
        # results are staged per input slot and collected one step late: with a single slot step s+1 would overwrite
        # (and re-record the event of) step s before it has been handed out
        if self.n_buf < max(2, self.n_lanes):
            raise ValueError("DecodePipeline needs at least two input buffers and at least one per lane")
        self.lanes = [torch.cuda.Stream(self.device) for _ in range(self.n_lanes)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self._bufs: List[Optional[Tuple[torch.Tensor, ...]]] = [None] * self.n_buf
        self._pinned: List[Optional[Tuple[torch.Tensor, ...]]] = [None] * self.n_buf
        self._ready = [torch.cuda.Event() for _ in range(self.n_buf)]
        self._freed = [torch.cuda.Event() for _ in range(self.n_buf)]
        self._done = [torch.cuda.Event() for _ in range(self.n_buf)]

    # ------------------------------------------------------------------ pieces
    def _stage(self, slot: int, host_batch: Sequence[torch.Tensor]):
        """H2D copies of one batch into input buffer `slot` on the copy stream."""
        bufs = self._bufs[slot]
        if bufs is None or any(b.shape != h.shape or b.dtype != h.dtype for b, h in zip(bufs, host_batch)):
            bufs = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_batch)
            for t in bufs:                                      # written on the copy stream, read on the lanes: a buffer
                t.record_stream(self.copy_stream)               # dropped later (shape change) must not be recycled early
                for ln in self.lanes:
                    t.record_stream(ln)
            self._bufs[slot] = bufs
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._freed[slot])      # the decode that last read this buffer has finished
            for d_t, h_t in zip(bufs, host_batch):
                d_t.copy_(h_t, non_blocking=True)
            self._ready[slot].record(self.copy_stream)

    def _decode(self, step: int, slot: int):
        lane_id = step % self.n_lanes
        lane, model = self.lanes[lane_id], self.models[lane_id]
        lane.wait_event(self._ready[slot])
        with torch.cuda.stream(lane):
            fn = model.beam_search_v_indexed if self.indexed else model.beam_search_v
            (words, gates), (lpw, lpg) = fn(self._bufs[slot], self.eos_idxs, self.beam_size, self.out_size, gt=self.gt)
            if self.post is not None:
                words = self.post(words)
            self._freed[slot].record(lane)
            outs = (words, gates, lpw, lpg)
            pin = self._pinned[slot]
            if pin is None or any(p.shape != o.shape for p, o in zip(pin, outs)):
                pin = tuple(torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs)
                self._pinned[slot] = pin
            for p, o in zip(pin, outs):
                p.copy_(o, non_blocking=True)                   # device -> host read of the step's result
            self._done[slot].record(lane)

    def _collect(self, slot: int):
        self._done[slot].synchronize()
        return tuple(p.clone() for p in self._pinned[slot])

    # ------------------------------------------------------------------ the loop
    def run(self, host_batches: Iterable[Sequence[torch.Tensor]]) -> Iterator[Tuple[torch.Tensor, ...]]:
        """host_batches yields tuples of pinned host tensors: (detections, det_seqs, verbs) or, with indexed=True,
        (detections, slot_index, verbs).  Yields (words, gates, lp_words, lp_gates) host tensors per batch, in order."""
        cur = torch.cuda.current_stream(self.device)
        for ev in self._freed:
            ev.record(cur)
        it = iter(host_batches)
        ahead = self.n_buf - self.n_lanes + 1       # batches staged before the decode that needs them is enqueued
        staged = 0
        pending: List[int] = []                     # slots of decodes enqueued but not yet handed out
        exhausted = False

        def stage_next():
            nonlocal staged, exhausted
            if exhausted:
                return False
            try:
                batch = next(it)
            except StopIteration:
                exhausted = True
                return False
            self._stage(staged % self.n_buf, batch)
            staged += 1
            return True

        for _ in range(ahead):
            stage_next()
        step = 0
        while step < staged:
            slot = step % self.n_buf
            self._decode(step, slot)
            pending.append(slot)
            step += 1
            stage_next()
            if len(pending) > 1:                    # hand out the previous step's result while this one decodes
                yield self._collect(pending.pop(0))
        while pending:
            yield self._collect(pending.pop(0))
