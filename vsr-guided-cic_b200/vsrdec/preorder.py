"""Eval pre-step: the per-caption role ordering of coco_scripts/eval_coco.py:127-237, batched.

(The integer bookkeeping exists twice: as the Python functions of this module — the readable statement, used with
`RoleOrderer(native=False)` and as the test twin — and as host code in libvsrdec, `vsr_preorder_*` (csrc/preorder.cu), which
`RoleOrderer` uses by default: at a thousand captions per call the Python form costs more than the device work it feeds.)

The reference walks every caption of a batch in Python and, per (caption, verb), calls `S_SSP.generate` on the GPU with batch 1,
per repeated role calls `SinkhornNet` with batch 1, copies the matrix to the host for munkres, and finally permutes the caption's
(10, R, F) slot tiles with an `np.dot` against a permutation matrix.  Here the integer bookkeeping stays on the host (it is a few
comparisons per caption), but

  * every (caption, verb) problem of the batch goes through ONE `S_SSP.generate_batch` call (csrc/sort.cu),
  * every repeated role of the batch goes through ONE `SinkhornNet.assign` call (csrc/ssp.cu: MLP + Sinkhorn + Hungarian),
  * the permutation is returned as slot indices: `src_slot (C, L)` — row j of the re-ordered caption is slot src_slot[c, j] —
    which `permute_slot_index` applies to the index-form input of `beam_search_v_indexed` (a (C, L, R) int32 gather) and
    `permute_slot_tiles` to the reference's materialised tiles (a device gather instead of the reference's matmul).

`RoleOrderer.order(...)` mirrors the names of the eval loop: control_verb, det_seqs_v, det_seqs_sr, verb_list, seqs_perm."""
from typing import Dict, List, Sequence, Tuple

import ctypes

import numpy as np
import torch


def merge_verb_ranks(first: Sequence[int], second: Sequence[int]) -> List[int]:
    """Slot order of a further verb merged into the order so far (utils/tools.py:35-71, `verb_rank_merge`): slots in both lists
    keep the first list's relative order (the second list is re-labelled to agree), and every slot only the second list has is
    inserted in front of the nearest shared slot on its right in the second list, or appended when there is none."""
    first, second = [int(x) for x in first], [int(x) for x in second]
    shared = [s for s in first if s in second]
    where = [second.index(s) for s in shared]
    if where != sorted(where):
        for s, p in zip(shared, sorted(where)):
            second[p] = s
    anchor, nearest = {}, None
    for s in reversed(second):
        if s in shared:
            nearest = s
        else:
            anchor[s] = nearest
    merged = list(first)
    for s in second:
        if s in shared:
            continue
        if anchor[s] is None:
            merged.append(s)
        elif anchor[s] in merged:
            merged.insert(merged.index(anchor[s]), s)
    return merged


def roles_of_verb(verb: int, det_seqs_v: np.ndarray, det_seqs_sr: np.ndarray, limit: int = 10):
    """(eval_coco.py:152-169) distinct roles of `verb` in first-seen order (at most `limit`; once that many are known every later
    match is ignored), the slots holding each role, and the roles held by several slots."""
    roles: List[int] = []
    slots: Dict[int, List[int]] = {}
    repeated: List[int] = []
    js, ks = np.nonzero(det_seqs_v == verb)           # row-major: slot by slot, then the slot's verb columns
    for j, k in zip(js.tolist(), ks.tolist()):
        if len(roles) >= limit:
            break
        sr = int(det_seqs_sr[j, k])
        if sr not in slots:
            slots[sr] = [j]
            roles.append(sr)
        else:
            slots[sr].append(j)
            if sr not in repeated:
                repeated.append(sr)
    return roles, slots, repeated


def permutation_from_rank(final_rank: Sequence[int], n_slots: int, slot_valid: Sequence[bool], verb_list: Sequence[float]):
    """(eval_coco.py:217-237) row j of the re-ordered caption is slot final_rank[j]; rows whose tile is empty are dropped and the
    tail repeats the last kept slot; the verb ids follow their slots un-compacted, -1 where nothing was placed."""
    placed = [int(r) for r in list(final_rank)[:n_slots]]
    kept = [r for r in placed if slot_valid[r]]
    src = kept + [kept[-1] if kept else -1] * (n_slots - len(kept))
    verbs = [float(verb_list[r]) for r in placed] + [-1.0] * (n_slots - len(placed))
    return src, verbs


def problems_of_batch(control_verb: np.ndarray, det_seqs_v: np.ndarray, det_seqs_sr: np.ndarray, limit: int = 10):
    """Every (caption, verb) problem of a batch, as `roles_of_verb` finds them one by one, from ONE vectorised match search:
    -> list of (caption, verb, roles, slots-by-role, repeated roles) in (caption, verb position) order; verbs after a
    caption's first 0 are ignored and verbs without a role are skipped (eval_coco.py:149-171)."""
    live = np.cumprod(control_verb != 0, axis=1).astype(bool)                       # (C, max_verb)
    hit = (det_seqs_v[:, None, :, :] == control_verb[:, :, None, None]) & live[:, :, None, None]   # (C, max_verb, L, max_verb)
    cs, vs, js, ks = np.nonzero(hit)                                               # row-major: caption, verb position, slot, column
    sr_all = det_seqs_sr[cs, js, ks].astype(np.int64).tolist()
    problems, cur = [], None
    for c, v, j, sr in zip(cs.tolist(), vs.tolist(), js.tolist(), sr_all):
        if cur is None or cur[0] != c or cur[1] != v:
            cur = [c, v, [], {}, []]
            problems.append(cur)
        roles, slots, repeated = cur[2], cur[3], cur[4]
        if len(roles) >= limit:
            continue
        if sr not in slots:
            slots[sr] = [j]
            roles.append(sr)
        else:
            slots[sr].append(j)
            if sr not in repeated:
                repeated.append(sr)
    return [(c, int(control_verb[c, v]), roles, slots, repeated) for c, v, roles, slots, repeated in problems]


class RoleOrderer:
    def __init__(self, sort_net, sinkhorn_net, sinkhorn_len: int = 10, fixed_len: int = 10, native: bool = True):
        """sort_net: models.S_SSP, sinkhorn_net: models.SinkhornNet(sinkhorn_len, ...), both on the CUDA device
        (eval_coco.py:94-103).  native: host bookkeeping in libvsrdec (vsr_preorder_*) instead of this module's Python."""
        self.sort_net, self.sinkhorn_net = sort_net, sinkhorn_net
        self.sinkhorn_len, self.fixed_len = int(sinkhorn_len), int(fixed_len)
        self.native = bool(native)

    # ------------------------------------------------------------------ native host path (csrc/preorder.cu)
    @staticmethod
    def _i64(x):
        a = x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        return np.ascontiguousarray(a, dtype=np.int64)

    def _native_begin(self, control_verb, det_seqs_v, det_seqs_sr, verb_list, seqs_perm, slot_valid):
        from . import _lib
        lib = _lib.load_library()
        cv, dv, ds = self._i64(control_verb), self._i64(det_seqs_v), self._i64(det_seqs_sr)
        C, n_verb, L = cv.shape[0], cv.shape[1], dv.shape[1]
        h, P, n_rep, max_roles = _lib.c_vp(), _lib.c_i32(), _lib.c_i32(), _lib.c_i32()
        ml = self.sort_net.max_len
        _lib.check(lib, lib.vsr_preorder_begin(cv.ctypes.data, dv.ctypes.data, ds.ctypes.data, C, n_verb, L, ml, self.sinkhorn_len,
                                               self.fixed_len, ctypes.byref(h), ctypes.byref(P), ctypes.byref(n_rep),
                                               ctypes.byref(max_roles)))
        P, n_rep = P.value, n_rep.value
        pred = assign = None
        try:
            if P:
                verbs = np.empty(P, dtype=np.int64)
                roles = np.empty((P, ml), dtype=np.int64)
                counts = np.empty(P, dtype=np.int32)
                gather = np.empty((n_rep, self.sinkhorn_len), dtype=np.int64)
                _lib.check(lib, lib.vsr_preorder_fill(h, verbs.ctypes.data, roles.ctypes.data, ml, counts.ctypes.data,
                                                      gather.ctypes.data if n_rep else None))
                dev = seqs_perm.device
                pred, _ = self.sort_net.generate_batch(torch.from_numpy(verbs).to(dev), torch.from_numpy(roles).to(dev),
                                                       counts=counts.tolist())
                if n_rep:
                    g = torch.from_numpy(gather).to(dev)
                    rows = seqs_perm.reshape(-1, seqs_perm.shape[-1]).float()
                    seq = rows[g.clamp(min=0)] * (g >= 0).unsqueeze(-1).to(rows.dtype)
                    _, assign = self.sinkhorn_net.assign(seq.contiguous())
        except Exception:
            lib.vsr_preorder_free(h)
            raise
        return ("native", h, C, pred, assign, verb_list, slot_valid)

    def _native_end(self, state):
        from . import _lib
        lib = _lib.load_library()
        _, h, C, pred, assign, verb_list, slot_valid = state
        pred_np = np.ascontiguousarray(pred.cpu().numpy(), dtype=np.int64) if pred is not None else None
        asg_np = np.ascontiguousarray(assign.cpu().numpy(), dtype=np.int32) if assign is not None else None
        sv = slot_valid.cpu().numpy() if isinstance(slot_valid, torch.Tensor) else np.asarray(slot_valid)
        sv = np.ascontiguousarray(sv.astype(bool), dtype=np.uint8).reshape(C, self.fixed_len)
        vl = verb_list.cpu().numpy() if isinstance(verb_list, torch.Tensor) else np.asarray(verb_list)
        vl = np.ascontiguousarray(vl, dtype=np.float64).reshape(C, self.fixed_len)
        src = np.empty((C, self.fixed_len), dtype=np.int64)
        verbs = np.empty((C, self.fixed_len), dtype=np.float32)
        rc = lib.vsr_preorder_end(h, None if pred_np is None else pred_np.ctypes.data, 0 if pred_np is None else pred_np.shape[1],
                                  None if asg_np is None else asg_np.ctypes.data, sv.ctypes.data, vl.ctypes.data,
                                  src.ctypes.data, verbs.ctypes.data)
        if rc != 0:
            lib.vsr_preorder_free(h)          # the library frees the state only on success
        _lib.check(lib, rc)
        return torch.from_numpy(src), torch.from_numpy(verbs)

    # The work splits at the only point where the host needs device results: `begin` does the host search and ENQUEUES the two
    # device calls (on the current stream), `end` reads their results back and assembles the ranks.  A loop that calls
    # begin(batch i + 1) under a side stream, enqueues the decode of batch i, then end(...) hides the pre-step's device time
    # behind the decode.
    def ranks_begin(self, control_verb, det_seqs_v, det_seqs_sr, seqs_perm):
        """control_verb (C, max_verb), det_seqs_v / det_seqs_sr (C, fixed_len, max_verb): host arrays / tensors;
        seqs_perm (C, fixed_len, 2352): the concatenated (vis, txt, pos) rows of eval_coco.py:146, on the CUDA device."""
        cv = np.asarray(control_verb.cpu() if isinstance(control_verb, torch.Tensor) else control_verb)
        dv = np.asarray(det_seqs_v.cpu() if isinstance(det_seqs_v, torch.Tensor) else det_seqs_v)
        ds = np.asarray(det_seqs_sr.cpu() if isinstance(det_seqs_sr, torch.Tensor) else det_seqs_sr)
        C = cv.shape[0]
        dev = seqs_perm.device
        problems = problems_of_batch(cv, dv, ds)
        if not problems:
            return (C, problems, [], None, None)
        # ---- S level: the order of every problem's roles, one device call
        L = self.sort_net.max_len
        roles_np = np.zeros((len(problems), L), dtype=np.int64)
        for i, p in enumerate(problems):
            roles_np[i, :len(p[2])] = p[2]
        verbs_np = np.fromiter((p[1] for p in problems), dtype=np.int64, count=len(problems))
        pred, _ = self.sort_net.generate_batch(torch.from_numpy(verbs_np).to(dev), torch.from_numpy(roles_np).to(dev),
                                               counts=[len(p[2]) for p in problems])
        # ---- R level: the order of the slots of every repeated role, one device call
        rep = [(i, sr) for i, p in enumerate(problems) for sr in p[4]]
        assign = None
        if rep:
            N = self.sinkhorn_len
            gather = np.full((len(rep), N), -1, dtype=np.int64)
            for n, (i, sr) in enumerate(rep):
                locs = problems[i][3][sr][:N]
                gather[n, :len(locs)] = locs
                gather[n, :len(locs)] += problems[i][0] * self.fixed_len
            gather = torch.from_numpy(gather).to(dev)
            rows = seqs_perm.reshape(-1, seqs_perm.shape[-1]).float()
            seq = rows[gather.clamp(min=0)] * (gather >= 0).unsqueeze(-1).to(rows.dtype)      # zero rows pad a role's problem
            _, assign = self.sinkhorn_net.assign(seq.contiguous())
        return (C, problems, rep, pred, assign)

    def ranks_end(self, state) -> List[List[int]]:
        """final_rank of every caption (eval_coco.py:148-215) from the state `ranks_begin` returned."""
        C, problems, rep, pred, assign = state
        if not problems:
            return [[] for _ in range(C)]
        pred = pred.cpu().numpy().tolist()
        assign = assign.cpu().numpy() if assign is not None else None
        region_rank = {}
        for n, (i, sr) in enumerate(rep):
            locs = problems[i][3][sr]
            cols = assign[n, :len(locs)]
            region_rank[(i, sr)] = [locs[a] for a in np.argsort(cols, kind="stable").tolist()]
        # ---- assemble per verb, merge over a caption's verbs
        per_caption: List[List[List[int]]] = [[] for _ in range(C)]
        for i, (c, verb, roles, slots, repeated) in enumerate(problems):
            rank: List[int] = []
            for sr in pred[i]:
                if sr == 0:
                    break
                rank += region_rank[(i, sr)] if len(slots[sr]) != 1 else slots[sr]
            per_caption[c].append(rank)
        out = []
        for ranks in per_caption:
            final = ranks[0] if ranks else []
            for nxt in ranks[1:]:
                final = merge_verb_ranks(final, nxt)
            out.append([int(x) for x in final])
        return out

    def ranks(self, control_verb, det_seqs_v, det_seqs_sr, seqs_perm) -> List[List[int]]:
        return self.ranks_end(self.ranks_begin(control_verb, det_seqs_v, det_seqs_sr, seqs_perm))

    def order_begin(self, control_verb, det_seqs_v, det_seqs_sr, verb_list, seqs_perm, slot_valid):
        if self.native:
            return self._native_begin(control_verb, det_seqs_v, det_seqs_sr, verb_list, seqs_perm, slot_valid)
        return (self.ranks_begin(control_verb, det_seqs_v, det_seqs_sr, seqs_perm), verb_list, slot_valid)

    def order_end(self, state) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (src_slot (C, fixed_len) long, verbs (C, fixed_len) float), host tensors."""
        if state[0] == "native":
            return self._native_end(state)
        rstate, verb_list, slot_valid = state
        ranks = self.ranks_end(rstate)
        sv = np.asarray(slot_valid.cpu() if isinstance(slot_valid, torch.Tensor) else slot_valid).astype(bool).tolist()
        vl = np.asarray(verb_list.cpu() if isinstance(verb_list, torch.Tensor) else verb_list).reshape(len(sv), -1).tolist()
        src = np.empty((len(ranks), self.fixed_len), dtype=np.int64)
        verbs = np.empty((len(ranks), self.fixed_len), dtype=np.float32)
        for c, rank in enumerate(ranks):
            src[c], verbs[c] = permutation_from_rank(rank, self.fixed_len, sv[c], vl[c])
        return torch.from_numpy(src), torch.from_numpy(verbs)

    def order(self, control_verb, det_seqs_v, det_seqs_sr, verb_list, seqs_perm, slot_valid) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (src_slot (C, fixed_len) long, verbs (C, fixed_len) float), host tensors.  slot_valid (C, fixed_len) bool: slots whose
        tile is not empty (`np.sum(tile) != 0`, eval_coco.py:226).  verb_list (C, fixed_len[, 1])."""
        return self.order_end(self.order_begin(control_verb, det_seqs_v, det_seqs_sr, verb_list, seqs_perm, slot_valid))


def permute_slot_index(slot_index: torch.Tensor, src_slot: torch.Tensor) -> torch.Tensor:
    """Index-form slots (C, L, R) int32 re-ordered: out[c, j] = slot_index[c, src_slot[c, j]], all padding (-1) where
    src_slot is -1.  The result feeds beam_search_v_indexed / vsr_prologue_indexed."""
    src = src_slot.to(slot_index.device)
    out = torch.gather(slot_index, 1, src.clamp(min=0).unsqueeze(-1).expand(-1, -1, slot_index.size(2)))
    return torch.where((src >= 0).unsqueeze(-1), out, torch.full_like(out, -1)).contiguous()


def permute_slot_tiles(det_seqs_all: torch.Tensor, src_slot: torch.Tensor) -> torch.Tensor:
    """The reference's materialised (C, L, R, F) slot tiles re-ordered the same way (`det_seqs_recons`, eval_coco.py:222-231)."""
    src = src_slot.to(det_seqs_all.device)
    idx = src.clamp(min=0).view(src.size(0), src.size(1), 1, 1).expand(-1, -1, det_seqs_all.size(2), det_seqs_all.size(3))
    out = torch.gather(det_seqs_all, 1, idx)
    return out * (src >= 0).view(src.size(0), src.size(1), 1, 1).to(out.dtype)
