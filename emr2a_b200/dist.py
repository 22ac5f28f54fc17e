"""Row-sharded database across GPUs: one process per GPU, ``torch.distributed`` (NCCL over
NVLink/NVSwitch) for the single exchange step of the path -- an all-gather of each rank's
local Top-K keys (Q x K x 8 bytes), merged by the K3 kernel.  Global row indices travel
inside the keys, and ties break on the global index, so results are bit-identical for any
GPU count."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_rows: int, rank: int, world: int, align: int = 256) -> Tuple[int, int]:
    """Contiguous, tile-aligned row range [lo, hi) of ``rank``; the last rank takes the tail."""
    per = (n_rows + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(rank * per, n_rows)
    hi = min(lo + per, n_rows)
    return lo, hi


def gather_keys(local_keys: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather [Q, K] packed keys -> [world, Q, K].  Works for CUDA tensors (NCCL) and for
    CPU tensors (gloo; used by the CPU tests of this plumbing)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_keys.unsqueeze(0)
    local_keys = local_keys.contiguous()
    out = torch.empty((world,) + tuple(local_keys.shape), dtype=local_keys.dtype, device=local_keys.device)
    if local_keys.is_cuda:
        dist.all_gather_into_tensor(out.view(-1), local_keys.view(-1), group=group)
    else:
        parts = [out[r] for r in range(world)]
        dist.all_gather(parts, local_keys, group=group)
    return out


def sharded_search_and_vote(eng, db_segs_local: Sequence, q_segs: Sequence, db_labels_global, q_labels,
                            n_classes: int, k: int, row_offset: int, db_flags: int, q_flags: int,
                            q_weights=(1.0, 1.0), k_list=(1, 3, 5), precision: str = "auto",
                            q_fold=None, db_fold_local=None, q_group=None, n_groups: int = 1,
                            want_lists: bool = True, timers: Optional[dict] = None,
                            defer_status: bool = False) -> Dict[str, torch.Tensor]:
    """Each rank: K1 on its database shard and on the (replicated) queries, local K2 with
    ``idx_base = row_offset``; all-gather keys; K3 merge; K4 vote on the merged lists."""
    n_q = int(q_segs[0].shape[0])
    n_db = int(db_segs_local[0].shape[0])
    dim = sum(int(s.shape[1]) for s in db_segs_local if s is not None)
    world = dist.get_world_size() if dist.is_initialized() else 1
    prec = eng.pick_precision(n_q, max(n_db, 1) * world, dim, k, precision)
    db = eng.prepare(db_segs_local[0], db_segs_local[1] if len(db_segs_local) > 1 else None, 1.0, 1.0, db_flags, prec)
    qs = eng.prepare(q_segs[0], q_segs[1] if len(q_segs) > 1 else None, q_weights[0], q_weights[1], q_flags, prec)
    if timers is not None:
        timers["k2_start"].record()
    keys = eng.topk_search(qs, db, k, prec, q_fold=q_fold, db_fold=db_fold_local, idx_base=row_offset)
    if timers is not None:
        timers["k2_end"].record()
    # optimistic execution: the rescore status (did any rank overflow its exact re-scan list?) travels
    # with the keys in the same all-gather and is only looked at after the vote, so the step has a
    # single host synchronisation at its very end
    status = eng.pop_status_tensor() if prec == "rescore" else None
    if world > 1:
        payload = keys.reshape(-1)
        if status is not None:
            payload = torch.cat([payload, status.to(torch.int64)])
        allp = gather_keys(payload.unsqueeze(0)).squeeze(1)              # [world, Q*K (+4)]
        allk = allp[:, :keys.numel()].reshape(world, *keys.shape)
        if status is not None:
            status = allp[:, keys.numel():].to(torch.int32).max(dim=0).values
        keys = eng.topk_merge(allk, k)
    res = eng.vote_metrics(keys, db_labels_global, q_labels, n_classes, k_list=k_list, q_group=q_group,
                           n_groups=n_groups, want_lists=want_lists)
    res["keys"] = keys
    res["precision"] = prec
    if status is not None and defer_status:
        res["status"] = status          # device int32[4]: the caller checks [1] == 0 (see Engine.check_deferred)
    elif status is not None:
        st = status.cpu()
        if int(st[1]):            # some rank could not verify within its re-scan capacity: redo with the 3-pass arm
            return sharded_search_and_vote(eng, db_segs_local, q_segs, db_labels_global, q_labels, n_classes, k,
                                           row_offset, db_flags, q_flags, q_weights, k_list, "bf16x3", q_fold,
                                           db_fold_local, q_group, n_groups, want_lists, timers)
        res["unverified"] = int(st[0])
    return res


def sharded_cv_search_and_vote(eng, segs_local: Sequence, labels_global, folds_global, n_classes: int, k: int,
                               row_offset: int, flags: int, q_weights=(1.0, 1.0), k_list=(1, 3, 5),
                               precision: str = "auto", n_folds: Optional[int] = None, q_block: int = 262144,
                               fold_sorted: bool = True, want_lists: bool = False,
                               full_segs: Optional[Sequence] = None) -> Dict[str, torch.Tensor]:
    """The whole CV loop over a row-sharded cohort: EVERY case is a query against the cases of the OTHER folds
    (utils/cv_evaluator.py:349-376 builds exactly these train/test pairs, one fold at a time), database rows
    sharded over the ranks (``segs_local`` = this rank's rows ``[row_offset, row_offset + n_local)`` of every
    modality; shards as ``shard_range`` cuts them: equal, 256-aligned, the last one shorter).

    Every rank needs every case as a query, so the raw rows are all-gathered ONCE over NVLink; after that each
    rank searches all query blocks against its own shard without communication, and only the local Top-K keys
    are exchanged (one all-gather + K3 merge per block) before K4 votes with per-fold counters.  A collective per
    query block would make ranks whose shard lies in the block's own fold (nothing to do: all tiles skipped) wait
    for the others.  ``fold_sorted`` promises ``folds_global`` is non-decreasing (rows in fold order) so whole
    tiles of a single fold are skipped; pass False for arbitrary fold vectors (per-element mask only).

    ``full_segs``: the rows of ALL cases, if this rank already holds them (then nothing is all-gathered).

    Returns per-fold counters (``hit_counts [F, nk]``, ``vote_counts [F, 3]``, ``confusion [F, 2, C, C]``,
    ``group_sizes [F]``), ``unverified`` and, with ``want_lists``, the per-query outputs of all N cases
    (identical on every rank).  Results are bit-identical for every GPU count."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = eng.device
    mats = [eng._embedding(x)[0] for x in segs_local if x is not None]
    n_local = int(mats[0].shape[0])
    labels = eng.to_device(labels_global, torch.int32)
    folds = eng.to_device(folds_global, torch.uint8)
    n = int(labels.shape[0])
    if n_folds is None:
        n_folds = int(folds.max().item()) + 1 if n else 1
    dim = sum(int(m.shape[1]) for m in mats)
    prec = eng.pick_precision(n, n, dim, k, precision)
    # (1) one all-gather of the raw rows (equal-sized, zero-padded shards)
    if full_segs is not None:
        full = [eng._embedding(x)[0] for x in full_segs if x is not None]
    elif world > 1:
        per = shard_range(n, 0, world)[1]
        full = []
        for m in mats:
            buf = torch.empty((world * per, int(m.shape[1])), dtype=m.dtype, device=dev)
            src = m
            if n_local < per:
                src = torch.zeros((per, int(m.shape[1])), dtype=m.dtype, device=dev)
                src[:n_local] = m
            dist.all_gather_into_tensor(buf, src.contiguous())
            full.append(buf)
    else:
        full = mats
    # (2) local work, no communication
    seg1 = mats[1] if len(mats) > 1 else None
    db = eng.prepare(mats[0], seg1, 1.0, 1.0, flags, prec)
    db_fold = folds[row_offset:row_offset + n_local]
    local = []
    for b0 in range(0, n, q_block):
        b1 = min(b0 + q_block, n)
        qs = eng.prepare(full[0][b0:b1], full[1][b0:b1] if len(full) > 1 else None, q_weights[0], q_weights[1], flags, prec)
        local.append(eng.topk_search(qs, db, k, prec, q_fold=folds[b0:b1], db_fold=db_fold, fold_sorted=fold_sorted,
                                     idx_base=row_offset))
    # (3) exchange + merge + vote
    outs = []
    for i, b0 in enumerate(range(0, n, q_block)):
        b1 = min(b0 + q_block, n)
        keys = local[i]
        if world > 1:
            keys = eng.topk_merge(gather_keys(keys), k)
        outs.append(eng.vote_metrics(keys, labels, labels[b0:b1], n_classes, k_list=k_list, q_group=folds[b0:b1],
                                     n_groups=n_folds, per_query=want_lists, want_lists=want_lists))
        local[i] = None
    unverified = 0
    if prec == "rescore":
        st = eng.pop_status_tensor()
        if world > 1:
            st = st.to(torch.int64)
            dist.all_reduce(st, op=dist.ReduceOp.MAX)          # every rank must take the same branch below
        st = st.cpu()
        if int(st[1]):
            return sharded_cv_search_and_vote(eng, segs_local, labels_global, folds_global, n_classes, k, row_offset,
                                              flags, q_weights, k_list, "bf16x3", n_folds, q_block, fold_sorted, want_lists,
                                              full_segs)
        unverified = int(st[0])
    res: Dict[str, torch.Tensor] = {}
    for name in ("hit_counts", "vote_counts", "confusion", "group_sizes"):
        res[name] = torch.stack([o[name] for o in outs]).sum(dim=0)
    if want_lists:
        for name in ("top_idx", "top_scores", "top_labels", "pred_top1", "pred_vote", "pred_weighted"):
            res[name] = torch.cat([o[name] for o in outs])
    res["precision"] = prec
    res["unverified"] = unverified          # this rank's count (max over ranks when sharded)
    return res
