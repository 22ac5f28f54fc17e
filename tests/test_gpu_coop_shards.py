"""Cooperative row shards (include/emr2a.h: "K2 in stages"): emr2a_topk_filter -> MAX of the shards' K-th best filter
score -> emr2a_rescore_candidates -> merge -> emr2a_verify_merged -> emr2a_exact_rescan for flagged queries.

The shards of a multi-GPU run are emulated on ONE device (slices of the prepared database, one after the other; the two
collectives of emr2a_b200/dist.py:_cooperative_search_and_vote become a torch.maximum and a torch.stack), so the kernels
of every stage are exercised by the single-GPU test tier.  The result must be bit-identical to the unsharded searches
(bit for bit to the rescore arm, same selection as the exact fp32 arm), with and without the CV fold rule (utils/cv_evaluator.py:349-376)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from emr2a_b200.engine import get_engine
    return get_engine()


def _shard(op, lo, hi):
    from emr2a_b200.engine import Operand
    return Operand(n=hi - lo, dim=op.dim, f32=None if op.f32 is None else op.f32[lo:hi], hi=op.hi[lo:hi], stats=op.stats,
                   lazy=None if op.lazy is None else op.lazy.slice(lo, hi))


def _coop(eng, qs, db, bounds_rows, k, q_fold=None, db_fold=None, force_flag=None):
    """The protocol of dist._cooperative_search_and_vote over emulated shards.  Returns (keys, n_flagged)."""
    import torch
    shards = [(_shard(db, lo, hi), lo, None if db_fold is None else db_fold[lo:hi]) for lo, hi in bounds_rows]
    stage1 = [eng.topk_filter(qs, s, k, q_fold=q_fold, db_fold=f, idx_base=lo) for s, lo, f in shards]
    floor = torch.stack([kth for _, _, kth in stage1]).max(dim=0).values            # all-reduce MAX
    payloads = [eng.rescore_candidates(c, t, floor, qs, s, k, idx_base=lo) for (c, t, _), (s, lo, _) in zip(stage1, shards)]
    allp = torch.stack(payloads)                                                     # all-gather
    keys = eng.merge_payload(allp, qs.n, k)
    flags, status = eng.verify_merged(keys, allp, k)
    if force_flag is not None:
        flags[force_flag] = 1
    n_status = int(status.cpu()[0])
    idx = torch.nonzero(flags).squeeze(1).to(torch.int32)
    assert force_flag is not None or n_status == int(idx.numel())
    assert int(status.cpu()[1]) == (1 if n_status else 0)
    if int(idx.numel()):
        seed = keys.index_select(0, idx.long())           # the merged lists seed the cut of the filtered re-scan (as dist.py does)
        comp = torch.stack([eng.exact_rescan(qs, s, idx, k, idx_base=lo, q_fold=q_fold, db_fold=f, seed_keys=seed)
                            for s, lo, f in shards])
        keys.index_copy_(0, idx.long(), eng.topk_merge(comp, k))
    return keys, int(idx.numel())


def _assert_same_selection(eng, got, qo, dbo, k, q_fold=None, db_fold=None):
    """Against the exact fp32 arm (CUDA cores; its summation order differs from the re-scoring's, so scores agree to a
    few ulps, not bit for bit): scores within 2e-6, index rows identical wherever the fp32 arm's adjacent score gaps --
    including the one to the (K+1)-th row -- exceed 1e-5."""
    from emr2a_b200.engine import unpack_keys
    wide = eng.topk_search(qo, dbo, k + 1, "fp32", q_fold=q_fold, db_fold=db_fold)
    es, ei = unpack_keys(wide)
    gs, gi = unpack_keys(got)
    assert np.allclose(gs, es[:, :k], rtol=0, atol=2e-6)
    clear = (es[:, :-1] - es[:, 1:]).min(axis=1) > 1e-5
    assert clear.mean() > 0.5
    assert np.array_equal(gi[clear], ei[clear, :k])


def _case(seed, n, d, q, n_cls=3, dup=0):
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, n_cls, n)
    centers = rng.standard_normal((n_cls, d)).astype(np.float32)
    db = (centers[labels] * 0.5 + rng.standard_normal((n, d))).astype(np.float32)
    ql = rng.integers(0, n_cls, q)
    qs = (centers[ql] * 0.5 + rng.standard_normal((q, d))).astype(np.float32)
    if dup:                                   # near-duplicate neighbourhoods: the bound cannot verify these queries
        base = rng.standard_normal((1, d)).astype(np.float32)
        db[:dup] = base + 1e-4 * rng.standard_normal((dup, d)).astype(np.float32)
        qs[:8] = base + 1e-4 * rng.standard_normal((8, d)).astype(np.float32)
    return db, qs


@pytest.mark.parametrize("n,d,q,k,parts,defer", [(40000, 256, 700, 10, 4, False), (9000, 192, 130, 5, 3, True),
                                                 (40000, 256, 700, 10, 8, True), (3000, 64, 40, 10, 2, False)])
def test_cooperative_shards_equal_unsharded(eng, n, d, q, k, parts, defer):
    import torch
    from emr2a_b200 import native
    from emr2a_b200.dist import shard_range
    db, qs = _case(n + q, n, d, q)
    dbo = eng.prepare(db, None, 1.0, 1.0, native.NF_ROWNORM, "rescore", defer_f32=defer)      # deferred fp32 rows or not
    assert (dbo.f32 is None) == defer
    qo = eng.prepare(qs, None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    full = dbo if not defer else eng.prepare(db, None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    want = eng.topk_search(qo, full, k, "rescore")     # unsharded, materialised fp32 rows
    eng.consume_status()
    spans = [shard_range(n, r, parts) for r in range(parts)]
    got, n_flagged = _coop(eng, qo, dbo, spans, k)
    torch.cuda.synchronize()
    assert torch.equal(got, want)                      # bit-identical to the unsharded rescore arm
    _assert_same_selection(eng, got, qo, full, k)
    assert n_flagged <= q // 20          # merged verification: (almost) nothing left to re-scan


def test_cooperative_shards_with_fold_rule(eng):
    import torch
    from emr2a_b200 import native
    from emr2a_b200.dist import shard_range
    n, d, q, k, parts = 20000, 128, 512, 5, 4
    db, _ = _case(11, n, d, q)
    folds = (np.arange(n) * 5 // n).astype(np.uint8)
    pick = np.linspace(0, n - 1, q).astype(np.int64)            # queries ARE database rows: only the fold rule keeps them apart
    dbo = eng.prepare(db, None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    qo = eng.prepare(db[pick], None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    db_fold = eng.to_device(folds, torch.uint8)
    q_fold = db_fold[torch.from_numpy(pick).to(eng.device)]
    want = eng.topk_search(qo, dbo, k, "rescore", q_fold=q_fold, db_fold=db_fold)
    eng.consume_status()
    spans = [shard_range(n, r, parts) for r in range(parts)]
    got, _ = _coop(eng, qo, dbo, spans, k, q_fold=q_fold, db_fold=db_fold)
    assert torch.equal(got, want)
    _assert_same_selection(eng, got, qo, dbo, k, q_fold=q_fold, db_fold=db_fold)
    from emr2a_b200.engine import unpack_keys
    idx = unpack_keys(got)[1]
    assert not (folds[idx] == folds[pick][:, None]).any()


def test_flagged_queries_are_repaired_by_the_exact_rescan(eng):
    """Near-duplicate neighbourhoods (score gaps of ~1e-8 among 300 rows spread over all shards) defeat the bound: those
    queries must come back flagged, and the exact re-scan must make them equal to the unsharded rescore arm (which
    re-scans them too, with the same fp32 arithmetic).  A few verified queries are flagged by hand on top, which must
    not change them."""
    import torch
    from emr2a_b200 import native
    from emr2a_b200.dist import shard_range
    n, d, q, k, parts = 30000, 256, 200, 10, 4
    db, qs = _case(5, n, d, q, dup=300)
    perm = np.random.default_rng(1).permutation(n)              # spread the duplicates over the shards
    db = db[perm]
    dbo = eng.prepare(db, None, 1.0, 1.0, native.NF_ROWNORM, "rescore", defer_f32=True)
    qo = eng.prepare(qs, None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    full = eng.prepare(db, None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    want = eng.topk_search(qo, full, k, "rescore")
    assert eng.consume_status()[0] >= 8
    spans = [shard_range(n, r, parts) for r in range(parts)]
    got, n_flagged = _coop(eng, qo, dbo, spans, k, force_flag=torch.tensor([50, 51, 199], device=eng.device))
    assert n_flagged >= 8 + 3
    assert torch.equal(got, want)
    from emr2a_b200.engine import unpack_keys
    wide = unpack_keys(eng.topk_search(qo, full, k, "fp32"))[0]
    assert np.allclose(unpack_keys(got)[0], wide, rtol=0, atol=2e-6)


def test_filter_pads_short_lists_and_handles_tiny_shards(eng):
    """Shards with fewer rows than candidates / than K: empty slots are zero keys, kth = -inf, bound = -inf."""
    import torch
    from emr2a_b200 import native
    n, d, q, k = 700, 64, 33, 10
    db, qs = _case(2, n, d, q)
    dbo = eng.prepare(db, None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    qo = eng.prepare(qs, None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    want = eng.topk_search(qo, dbo, k, "rescore")
    eng.consume_status()
    spans = [(0, 6), (6, 256), (256, 256), (256, 700)]          # 6 rows (< K), 250 rows, EMPTY, 444 rows
    cand, tau, kth = eng.topk_filter(qo, _shard(dbo, 0, 6), k)
    assert torch.isinf(kth).all() and (kth < 0).all()
    assert (cand[:, 6:] == 0).all() and (cand[:, :6] != 0).all()
    got, _ = _coop(eng, qo, dbo, spans, k)
    assert torch.equal(got, want)
    _assert_same_selection(eng, got, qo, dbo, k)
