"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: shard ranges, the all-gather of packed
Top-K keys, and the property that merging per-shard Top-K lists by key equals the global Top-K
(so results cannot depend on the GPU count)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def _pack(scores, idx):
    b = scores.astype(np.float32).view(np.uint32).astype(np.uint64)
    o = np.where(b & np.uint64(0x80000000), (~b) & np.uint64(0xFFFFFFFF), b ^ np.uint64(0x80000000))
    return ((o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - idx.astype(np.uint64))).view(np.int64)


def _local_topk(scores, lo, k):
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return _pack(np.take_along_axis(scores, order, axis=1), order + lo)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emr2a_b200.dist import gather_keys, shard_range
    rng = np.random.default_rng(0)                      # same data on every rank
    n, n_q, k = 1000, 17, 5
    scores = rng.standard_normal((n_q, n)).astype(np.float32)
    scores[:, 700] = scores[:, 100]                     # ties across the shard boundary
    lo, hi = shard_range(n, rank, world, align=256)
    local = torch.from_numpy(_local_topk(scores[:, lo:hi], lo, k))
    allk = gather_keys(local).numpy()                   # [world, Q, K]
    flat = np.transpose(allk, (1, 0, 2)).reshape(n_q, -1).view(np.uint64)
    merged = np.sort(flat, axis=1)[:, ::-1][:, :k]          # what the K3 kernel computes on the GPU
    want = _local_topk(scores, 0, k).view(np.uint64)
    q.put((rank, bool(np.array_equal(merged, want)), (lo, hi)))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from emr2a_b200.dist import shard_range
    for n in (0, 1, 255, 256, 1000, 1_000_000, 10_000_001):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all(lo % 256 == 0 or lo == n for lo, _ in spans)


def test_gather_and_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [o[1] for o in out] == [True, True]
    assert out[0][2] == (0, 512) and out[1][2] == (512, 1000)


# ---------------------------------------------------------------------------------------------------------------
# dist.sharded_cv_search_and_vote (the multi-GPU all-queries CV, BASELINE config 5) under gloo with a numpy stand-in
# for the CUDA engine: checks the host-side plumbing -- shard padding and the all-gather of the raw rows, row
# offsets / global indices, fold slices per query block, key exchange and merge order, per-fold counters.
class _StubEngine:
    device = torch.device("cpu")
    launches = 0

    def _embedding(self, x):
        return torch.as_tensor(np.asarray(x), dtype=torch.float32), 0

    def to_device(self, x, dtype=None):
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        return t.to(dtype) if dtype is not None else t

    def pick_precision(self, q, n, d, k, requested="auto"):
        return "fp32"

    def prepare(self, s0, s1, w0, w1, flags, prec):
        mats = [np.asarray(s0, dtype=np.float32) * np.float32(w0)] + ([np.asarray(s1, dtype=np.float32) * np.float32(w1)] if s1 is not None else [])
        rows = np.concatenate(mats, axis=1)
        from emr2a_b200.engine import Operand
        rows = rows / (np.linalg.norm(rows, axis=1, keepdims=True) + 1e-8)
        return Operand(n=rows.shape[0], dim=rows.shape[1], f32=torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32)))

    def topk_search(self, qs, db, k, prec, q_fold=None, db_fold=None, fold_sorted=False, idx_base=0, row_ids=None):
        sc = (qs.f32.numpy().astype(np.float64) @ db.f32.numpy().astype(np.float64).T).astype(np.float32)
        if q_fold is not None:
            sc = np.where(q_fold.numpy()[:, None] == db_fold.numpy()[None, :], -np.inf, sc)
        kk = min(k, sc.shape[1])
        order = np.argsort(-sc, axis=1, kind="stable")[:, :kk]
        top = np.take_along_axis(sc, order, axis=1)
        keys = np.zeros((sc.shape[0], k), dtype=np.int64)
        gidx = order + idx_base if row_ids is None else row_ids.numpy().astype(np.int64)[order]
        keys[:, :kk] = np.where(np.isfinite(top), _pack(top, gidx), 0)
        return torch.from_numpy(keys)

    def topk_merge(self, parts, k):
        p, q, kin = parts.shape
        flat = parts.numpy().view(np.uint64).transpose(1, 0, 2).reshape(q, p * kin)
        return torch.from_numpy(np.ascontiguousarray(np.sort(flat, axis=1)[:, ::-1][:, :k]).view(np.int64))

    def vote_metrics(self, keys, db_labels, q_labels, n_classes, k_list=(1, 3, 5), q_group=None, n_groups=1,
                     per_query=True, want_lists=True, **_):
        u = keys.numpy().view(np.uint64)
        idx = np.where(u == 0, -1, (np.uint64(0xFFFFFFFF) - (u & np.uint64(0xFFFFFFFF))).astype(np.int64))
        lab = np.where(idx >= 0, db_labels.numpy()[np.clip(idx, 0, None)], -1)
        grp = q_group.numpy().astype(np.int64) if q_group is not None else np.zeros(len(u), dtype=np.int64)
        hit = np.zeros((n_groups, len(k_list)), dtype=np.int64)
        for j, kk in enumerate(k_list):
            np.add.at(hit[:, j], grp, (lab[:, :kk] == q_labels.numpy()[:, None]).any(axis=1))
        votes = np.zeros((n_groups, 3), dtype=np.int64)
        np.add.at(votes[:, 0], grp, lab[:, 0] == q_labels.numpy())
        sizes = np.bincount(grp, minlength=n_groups).astype(np.int64)
        res = {"hit_counts": torch.from_numpy(hit), "vote_counts": torch.from_numpy(votes),
               "confusion": torch.zeros((n_groups, 2, n_classes, n_classes), dtype=torch.int64),
               "group_sizes": torch.from_numpy(sizes)}
        if want_lists:
            z = torch.zeros((len(u),), dtype=torch.int32)
            res.update(top_idx=torch.from_numpy(idx), top_scores=torch.zeros(idx.shape), top_labels=torch.from_numpy(lab),
                       pred_top1=torch.from_numpy(lab[:, 0].astype(np.int32)), pred_vote=z, pred_weighted=z)
        return res


def _cv_case():
    rng = np.random.default_rng(5)
    n, d0, d1, n_folds = 1100, 12, 9, 5                   # 1100 rows over 2 ranks: shards of 768 and 332 (padding path)
    labels = rng.integers(0, 3, n).astype(np.int32)
    a = (rng.standard_normal((n, d0)) + labels[:, None]).astype(np.float32)
    b = rng.standard_normal((n, d1)).astype(np.float32)
    folds = (np.arange(n) * n_folds // n).astype(np.uint8)
    return a, b, labels, folds, n_folds


def _cv_run(world, rank, balanced=False):
    """contiguous shards (row_offset) or fold-balanced shards (row_ids: a slice of every fold per rank)"""
    from emr2a_b200.dist import fold_balanced_ranges, ranges_to_rows, shard_range, sharded_cv_search_and_vote
    a, b, labels, folds, n_folds = _cv_case()
    if balanced:
        rows = ranges_to_rows(fold_balanced_ranges(np.bincount(folds, minlength=n_folds), rank, world, align=64), "cpu")
        sel = rows.numpy()
        r = sharded_cv_search_and_vote(_StubEngine(), (a[sel], b[sel]), labels, folds, 3, 4, 0, 0, q_weights=(0.7, 0.3),
                                       k_list=(1, 3), precision="fp32", n_folds=n_folds, q_block=100, want_lists=True,
                                       row_ids=rows)
    else:
        lo, hi = shard_range(len(labels), rank, world)
        r = sharded_cv_search_and_vote(_StubEngine(), (a[lo:hi], b[lo:hi]), labels, folds, 3, 4, lo, 0, q_weights=(0.7, 0.3),
                                       k_list=(1, 3), precision="fp32", n_folds=n_folds, q_block=256, want_lists=True)
    return {k: r[k].numpy() for k in ("hit_counts", "vote_counts", "group_sizes", "top_idx", "top_labels")}


def _cv_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {"contiguous": _cv_run(world, rank), "balanced": _cv_run(world, rank, balanced=True)}
    q.put((rank, out))
    dist.destroy_process_group()


def test_fold_balanced_ranges_partition_every_fold():
    from emr2a_b200.dist import fold_balanced_ranges
    for counts in ([2_000_000] * 5, [10, 300, 7], [0, 5, 0], [1999999, 2000001, 2000000, 2000000, 2000000]):
        for world in (1, 2, 4, 8):
            per_rank = [fold_balanced_ranges(counts, r, world) for r in range(world)]
            start = 0
            for f, n_f in enumerate(counts):
                pieces = [pr[f] for pr in per_rank]
                assert pieces[0][0] == start and sum(c for _, c in pieces) == n_f
                for (a0, c0), (a1, _) in zip(pieces, pieces[1:]):
                    assert a0 + c0 == a1
                assert all((g0 - start) % 256 == 0 or c == 0 for g0, c in pieces)
                start += n_f
            sizes = [sum(c for _, c in pr) for pr in per_rank]
            if min(counts) >= 256 * world:
                assert max(sizes) - min(sizes) <= 256 * world * len(counts)     # shards of (nearly) equal size


def test_sharded_cv_world2_equals_single_process():
    single = _cv_run(1, 0)
    a, b, labels, folds, _ = _cv_case()
    assert not (folds[single["top_idx"]] == folds[:, None]).any()          # the CV rule holds in the stand-in too
    assert single["group_sizes"].sum() == len(labels)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_cv_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(np.array_equal(_cv_run(1, 0, balanced=True)[key], want) for key, want in single.items())
    for rank in (0, 1):
        for layout in ("contiguous", "balanced"):                                 # any sharding, same result
            for key, want in single.items():
                assert np.array_equal(outs[rank][layout][key], want), (rank, layout, key)
