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

    def defer_default(self, dim):
        return False

    def prepare(self, s0, s1, w0, w1, flags, prec, defer_f32=False):
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


# ---------------------------------------------------------------------------------------------------------------
# dist.sharded_search_and_vote, cooperative row shards ("K2 in stages", include/emr2a.h) under gloo: the two
# collectives (all-reduce MAX of the K-th best filter score, all-gather of exact keys + bounds in ONE buffer), the
# verification of the merged lists and the repair of flagged queries -- with a numpy stand-in for the staged kernels
# whose "filter" is the exact score plus a bounded perturbation (|s~ - s| <= E), as the bf16 filter is.
class _CoopStub(_StubEngine):
    E = np.float32(0.02)
    WIDTH = 12                                   # candidates per query and shard (64 on the GPU)

    def pick_precision(self, q, n, d, k, requested="auto"):
        return "rescore"

    def _scores(self, qs, db, q_fold, db_fold):
        sc = (qs.f32.numpy().astype(np.float64) @ db.f32.numpy().astype(np.float64).T).astype(np.float32)
        if q_fold is not None:
            sc = np.where(q_fold.numpy()[:, None] == db_fold.numpy()[None, :], -np.inf, sc)
        return sc

    def topk_filter(self, qs, db, k, q_fold=None, db_fold=None, fold_sorted=False, idx_base=0):
        sc = self._scores(qs, db, q_fold, db_fold)
        gidx = np.arange(sc.shape[1]) + idx_base
        noise = (((np.arange(sc.shape[0])[:, None] * 7919 + gidx[None, :] * 104729) % 2001) / 1000.0 - 1.0).astype(np.float32)
        approx = sc + self.E * noise                                              # deterministic, |error| <= E
        w = self.WIDTH
        order = np.argsort(-approx, axis=1, kind="stable")
        top = np.take_along_axis(approx, order, axis=1)
        cand = np.zeros((sc.shape[0], w), dtype=np.int64)
        kk = min(w, sc.shape[1])
        cand[:, :kk] = np.where(np.isfinite(top[:, :kk]), _pack(top[:, :kk], order[:, :kk] + idx_base), 0)
        tau = np.full(sc.shape[0], -np.inf, dtype=np.float32)                     # best filter score outside the list
        if sc.shape[1] > w:
            tau = top[:, w].astype(np.float32)
        kth = top[:, k - 1].astype(np.float32) if sc.shape[1] >= k else np.full(sc.shape[0], -np.inf, dtype=np.float32)
        kth = np.where(np.isfinite(kth), kth, -np.inf).astype(np.float32)
        self.launches += 3
        return torch.from_numpy(cand), torch.from_numpy(tau), torch.from_numpy(kth)

    def rescore_candidates(self, cand, tau, kth_floor, qs, db, k, idx_base=0):
        from emr2a_b200.engine import Engine
        Q = qs.n
        u = cand.numpy().view(np.uint64)
        bits = (u >> np.uint64(32)).astype(np.uint32)
        fb = np.where(bits & np.uint32(0x80000000), bits ^ np.uint32(0x80000000), ~bits).astype(np.uint32)
        approx = np.where(u == 0, -np.inf, fb.view(np.float32))
        idx = (np.uint64(0xFFFFFFFF) - (u & np.uint64(0xFFFFFFFF))).astype(np.int64) - idx_base
        local_kth = approx[:, k - 1] if approx.shape[1] >= k else np.full(Q, -np.inf)
        cut = np.maximum(local_kth, kth_floor.numpy()) - 2 * self.E
        exact = np.full(approx.shape, -np.inf, dtype=np.float32)
        qf, df = qs.f32.numpy().astype(np.float64), db.f32.numpy().astype(np.float64)
        self.rescored = 0
        for i in range(Q):
            for j in range(approx.shape[1]):
                if u[i, j] != 0 and approx[i, j] >= cut[i]:
                    exact[i, j] = np.float32(qf[i] @ df[idx[i, j]])
                    self.rescored += 1
        order = np.argsort(-exact, axis=1, kind="stable")[:, :k]
        top = np.take_along_axis(exact, order, axis=1)
        keys = np.where(np.isfinite(top), _pack(top, np.take_along_axis(idx, order, axis=1) + idx_base), 0)
        if keys.shape[1] < k:
            keys = np.concatenate([keys, np.zeros((Q, k - keys.shape[1]), dtype=np.int64)], axis=1)
        # ties on the exact score must order by index like the keys do: sort the packed keys themselves
        keys = np.sort(keys.view(np.uint64), axis=1)[:, ::-1].copy().view(np.int64)
        payload = torch.zeros((Q * k + (Q + 1) // 2,), dtype=torch.int64)
        kv, bv = Engine.split_payload(payload, Q, k)
        kv.copy_(torch.from_numpy(np.ascontiguousarray(keys)))
        bv.copy_(torch.from_numpy(np.where(np.isfinite(tau.numpy()), tau.numpy() + self.E, -np.inf).astype(np.float32)))
        self.launches += 1
        return payload

    def merge_payload(self, allp, Q, k):
        from emr2a_b200.engine import Engine
        return self.topk_merge(Engine.split_payload(allp, Q, k)[0].contiguous(), k)

    def verify_merged(self, keys, allp, k):
        from emr2a_b200.engine import Engine
        Q = keys.shape[0]
        bounds = Engine.split_payload(allp, Q, k)[1].numpy().max(axis=0)
        u = keys.numpy().view(np.uint64)[:, k - 1]
        bits = (u >> np.uint64(32)).astype(np.uint32)
        kth = np.where(bits & np.uint32(0x80000000), bits ^ np.uint32(0x80000000), ~bits).astype(np.uint32).view(np.float32)
        ok = np.isneginf(bounds) | ((u != 0) & (kth > bounds))
        n_bad = int((~ok).sum())
        return torch.from_numpy((~ok).astype(np.uint8)), torch.tensor([n_bad, int(n_bad > 0), 0, 0], dtype=torch.int32)

    def exact_rescan(self, qs, db, flag_list, k, idx_base=0, q_fold=None, db_fold=None, seed_keys=None):
        from emr2a_b200.engine import Operand
        sel = flag_list.long()
        sub = Operand(n=int(sel.numel()), dim=qs.dim, f32=qs.f32[sel])
        return _StubEngine.topk_search(self, sub, db, k, "fp32", q_fold=None if q_fold is None else q_fold[sel],
                                       db_fold=db_fold, idx_base=idx_base)


def _coop_case():
    rng = np.random.default_rng(21)
    n, d, n_q, k = 1500, 16, 64, 5
    labels = rng.integers(0, 3, n).astype(np.int32)
    db = (rng.standard_normal((n, d)) + labels[:, None]).astype(np.float32)
    ql = rng.integers(0, 3, n_q).astype(np.int32)
    qs = (rng.standard_normal((n_q, d)) + ql[:, None]).astype(np.float32)
    return db, labels, qs, ql, k


def _coop_run(world, rank, width, defer=False):
    from emr2a_b200.dist import shard_range, sharded_search_and_vote
    db, labels, qs, ql, k = _coop_case()
    lo, hi = shard_range(len(labels), rank, world)
    eng = _CoopStub()
    eng.WIDTH = width
    r = sharded_search_and_vote(eng, (db[lo:hi],), (qs,), torch.from_numpy(labels), torch.from_numpy(ql), 3, k, lo, 0, 0,
                                k_list=(1, 3), precision="auto", defer_status=defer)
    out = {"keys": r["keys"].numpy().copy(), "hit_counts": r["hit_counts"].numpy().copy(),
           "unverified": int(r.get("unverified", -1)), "rescored": getattr(eng, "rescored", None)}
    if defer:
        out["status"] = r["status"].numpy().copy()
    return out


def _coop_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {"wide": _coop_run(world, rank, 40), "narrow": _coop_run(world, rank, 7), "deferred": _coop_run(world, rank, 7, defer=True)}
    q.put((rank, out))
    dist.destroy_process_group()


def test_cooperative_shards_world2_equal_exact_search():
    db, labels, qs, ql, k = _coop_case()
    stub = _StubEngine()
    want = stub.topk_search(stub.prepare(qs, None, 1, 1, 0, "fp32"), stub.prepare(db, None, 1, 1, 0, "fp32"), k, "fp32").numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_coop_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank in (0, 1):
        wide, narrow, deferred = outs[rank]["wide"], outs[rank]["narrow"], outs[rank]["deferred"]
        assert np.array_equal(wide["keys"], want) and np.array_equal(narrow["keys"], want)
        assert wide["unverified"] == 0                      # 40 candidates per shard: every merged selection verifies
        assert narrow["unverified"] > 0                     # 7 candidates for K = 5 and E = 0.02: the repair path ran
        assert np.array_equal(wide["hit_counts"], narrow["hit_counts"])
        # deferred status: no repair inside the call, [1] tells the caller the step must not be used as it is
        assert deferred["status"][0] == narrow["unverified"] and deferred["status"][1] == 1
        # the exchange of the K-th best filter score cuts the re-scoring: fewer candidates than shards x lists hold
        assert wide["rescored"] < 64 * 40
    assert outs[0]["narrow"]["unverified"] == outs[1]["narrow"]["unverified"]


def test_spread_device_and_weighted_ranges(monkeypatch):
    from emr2a_b200.dist import spread_device, weighted_ranges
    monkeypatch.delenv("EMR2A_SPREAD_DEVICES", raising=False)
    assert [spread_device(r, 4, 8) for r in range(4)] == [0, 2, 4, 6]          # spread over both host bridges
    assert [spread_device(r, 2, 8) for r in range(2)] == [0, 4]
    assert [spread_device(r, 8, 8) for r in range(8)] == list(range(8))
    assert spread_device(0, 1, 8) == 0 and spread_device(2, 3, 8) == 2          # 8 % 3 != 0: left alone
    assert [spread_device(r, 4, 4) for r in range(4)] == [0, 1, 2, 3]
    monkeypatch.setenv("EMR2A_SPREAD_DEVICES", "0")
    assert [spread_device(r, 4, 8) for r in range(4)] == [0, 1, 2, 3]
    # shard sizes follow the weights, cover every row once, and stay tile-aligned
    rates = [23.3, 23.4, 23.3, 23.3, 35.5, 35.4, 35.6, 35.6]                    # the pool's 8-GPU boxes (tools/h2d_probe.py)
    spans = weighted_ranges(1_000_000, rates)
    assert spans[0][0] == 0 and spans[-1][1] == 1_000_000
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all(lo % 256 == 0 for lo, _ in spans)
    sizes = np.array([hi - lo for lo, hi in spans], dtype=np.float64)
    assert np.allclose(sizes / sizes.sum(), np.array(rates) / sum(rates), atol=3e-4)
    times = sizes / np.array(rates)
    assert times.max() / times.min() < 1.01                                       # all copies end together
    assert weighted_ranges(1000, [1, 1]) == [(0, 512), (512, 1000)]
    assert weighted_ranges(10, [0, 0, 0]) == [(0, 0), (0, 0), (0, 10)]
    assert weighted_ranges(0, [1, 2]) == [(0, 0), (0, 0)]


def test_plan_chunks_covers_rows_and_tapers():
    from emr2a_b200.engine import plan_chunks
    for n, c in [(98816, 32768), (1_000_000, 32768), (5000, 32768), (0, 32768), (2500, 1000), (1500, 1000), (1001, 1000),
                 (999, 1000), (300, 256), (7, 3), (1, 1), (125_184, 32768)]:
        ch = plan_chunks(n, c)
        assert (not ch and n == 0) or (ch[0][0] == 0 and ch[-1][1] == n)
        assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        assert all(0 < hi - lo <= c for lo, hi in ch)                   # every piece fits a chunk-sized staging slot
    sizes = [hi - lo for lo, hi in plan_chunks(1_000_000, 32768)]
    assert sizes[0] == 32768 and sizes[-1] <= 8192 and sorted(sizes, reverse=True) == sizes


def _gather_rows_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emr2a_b200.dist import gather_host_rows
    rng = np.random.default_rng(3)                       # same host matrices on every rank
    mats = [torch.from_numpy(rng.standard_normal((101, 7)).astype(np.float32)),        # 101 rows: ragged last slice
            torch.from_numpy(rng.standard_normal((101, 3)).astype(np.float32))]
    got, copied = gather_host_rows(mats, torch.device("cpu"))
    q.put((rank, all(torch.equal(g, m) for g, m in zip(got, mats)), copied))
    dist.destroy_process_group()


def test_gather_host_rows_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_rows_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [o[1] for o in out] == [True, True]
    assert out[0][2] == 51 * 10 * 4 and out[1][2] == 50 * 10 * 4          # each rank copied only its slice from the host
