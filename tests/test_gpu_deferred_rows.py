"""Deferred fp32 rows (include/emr2a.h: emr2a_lazy_rows; csrc/row_math.cuh): with the rescore arithmetic K1 does not
write the fp32 copy of the DATABASE -- it records the divisors it used per row, and the re-scoring / exact re-scan
kernels re-create the elements they need from the raw rows.  The claim is bit-identity: x -> RN(x / n_seg) -> * w ->
RN(. / n_row) are the same correctly rounded operations on the same inputs as in K1 (utils/cv_evaluator.py:95-105 in
numpy), so every key (score bits and index) must equal the one obtained from materialised rows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from emr2a_b200.engine import get_engine
    return get_engine()


def _rows(rng, n, d, n_cls=3, scale=1.0):
    lab = rng.integers(0, n_cls, n)
    cen = rng.standard_normal((n_cls, d)).astype(np.float32)
    return ((cen[lab] * 0.5 + rng.standard_normal((n, d))) * scale).astype(np.float32)


CASES = [
    # n, d0, d1, q, k, flags, (w0, w1), dtype
    (30000, 512, 512, 300, 10, "SEGNORM|ROWNORM", (1.0, 1.0), "f32"),        # C2 layout
    (20000, 256, 0, 200, 5, "ROWNORM", (1.0, 1.0), "f32"),                   # one segment
    (20000, 128, 384, 260, 10, "SEGNORM", (0.75, 0.25), "f32"),              # late-fusion layout, weights on the rows
    (12000, 64, 8, 150, 10, "SEGNORM|ROWNORM", (1.0, 2.0), "f32"),           # narrow rows (small register path)
    (3000, 4096, 1024, 140, 10, "SEGNORM|ROWNORM", (1.0, 1.0), "bf16"),      # C4 layout: bf16 inputs, block-per-row K1
    (6000, 1536, 512, 140, 5, "ROWNORM", (1.0, 1.0), "f32"),                 # 2048 columns: widest register-cached K1
    (8000, 96, 32, 130, 10, "ZERO_GUARD", (1.0, 1.0), "f32"),                # utils/common.py:4-8, zero rows stay zero
]


def _flags(native, names):
    out = 0
    for nm in names.split("|"):
        out |= getattr(native, "NF_" + nm)
    return out


@pytest.mark.parametrize("n,d0,d1,q,k,flag_names,w,dtype", CASES)
def test_deferred_rows_give_identical_keys(eng, n, d0, d1, q, k, flag_names, w, dtype):
    import torch
    from emr2a_b200 import native
    rng = np.random.default_rng(n + d0 + d1)
    flags = _flags(native, flag_names)
    a = torch.from_numpy(_rows(rng, n, d0, scale=3.0)).to(eng.device)
    b = torch.from_numpy(_rows(rng, n, d1, scale=0.2)).to(eng.device) if d1 else None
    qa = torch.from_numpy(_rows(rng, q, d0, scale=3.0)).to(eng.device)
    qb = torch.from_numpy(_rows(rng, q, d1, scale=0.2)).to(eng.device) if d1 else None
    if flag_names == "ZERO_GUARD":
        a[17] = 0
        b[17] = 0
    if dtype == "bf16":
        a, b, qa, qb = a.bfloat16(), b.bfloat16(), qa.bfloat16(), qb.bfloat16()
    full = eng.prepare(a, b, w[0], w[1], flags, "rescore")
    lazy = eng.prepare(a, b, w[0], w[1], flags, "rescore", defer_f32=True)
    assert full.f32 is not None and lazy.f32 is None and lazy.lazy is not None
    assert torch.equal(full.hi, lazy.hi) and torch.equal(full.stats, lazy.stats)
    qs = eng.prepare(qa, qb, 1.0, 1.0, flags, "rescore")
    want = eng.topk_search(qs, full, k, "rescore")
    st_full = eng.consume_status()
    got = eng.topk_search(qs, lazy, k, "rescore")
    st_lazy = eng.consume_status()
    assert torch.equal(got, want)
    assert st_full == st_lazy
    # every row, not only the candidates: the exact re-scan of ALL queries walks the whole database
    every = torch.arange(q, dtype=torch.int32, device=eng.device)[:64]
    rescanned = eng.exact_rescan(qs, lazy, every, k)            # unfiltered: every row re-created and scored exactly
    assert torch.equal(rescanned, eng.exact_rescan(qs, full, every, k))
    assert torch.equal(rescanned, want[:64])                     # and it finds what the verified search found
    # filtered re-scan (the bf16 plane is streamed, exact scoring on demand), seeded with lists that are deliberately
    # too low -- the exact Top-K of the FIRST HALF of the database -- so that the scan has better rows to find
    half = n // 2
    from emr2a_b200.engine import Operand
    first = Operand(n=half, dim=full.dim, f32=full.f32[:half], hi=full.hi[:half], stats=full.stats)
    seed = eng.exact_rescan(qs, first, every, k)
    for op in (full, lazy):
        assert torch.equal(eng.exact_rescan(qs, op, every, k, seed_keys=seed), rescanned)
    assert torch.equal(eng.exact_rescan(qs, lazy, every, k, seed_keys=rescanned), rescanned)   # a perfect seed changes nothing


def test_deferred_rows_recreate_every_element(eng):
    """One-hot queries turn the exact re-scan into a read-out of single elements: score(query e_j, row r) = out[r, j]
    exactly (one product by 1.0, the rest adds zeros), so the Top-K keys of query j carry the K largest values of
    column j -- the same bits from deferred and from materialised rows, and equal to the fp32 rows K1 writes."""
    import torch
    from emr2a_b200 import native
    from emr2a_b200.engine import Operand, unpack_keys
    rng = np.random.default_rng(8)
    n, d0, d1, k = 5000, 40, 24, 10
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    a = torch.from_numpy(_rows(rng, n, d0, scale=5.0)).to(eng.device)
    b = torch.from_numpy(_rows(rng, n, d1, scale=0.3)).to(eng.device)
    full = eng.prepare(a, b, 0.6, 1.7, flags, "rescore")
    lazy = eng.prepare(a, b, 0.6, 1.7, flags, "rescore", defer_f32=True)
    eye = torch.eye(d0 + d1, dtype=torch.float32, device=eng.device)
    qs = Operand(n=d0 + d1, dim=d0 + d1, f32=eye)
    cols = torch.arange(d0 + d1, dtype=torch.int32, device=eng.device)
    got = eng.exact_rescan(qs, lazy, cols, k)
    assert torch.equal(got, eng.exact_rescan(qs, full, cols, k))
    scores, idx = unpack_keys(got)
    rows = full.f32.cpu().numpy()
    for j in range(d0 + d1):
        assert np.array_equal(scores[j], rows[idx[j], j])                      # the very fp32 values K1 stores
        assert np.array_equal(np.sort(rows[:, j])[::-1][:k], scores[j])        # and they are the column's K largest


def test_deferred_rows_through_rescan_and_fold_rule(eng):
    """Near-duplicate neighbourhoods force the exact re-scan; the CV fold rule masks rows -- both on deferred rows."""
    import torch
    from emr2a_b200 import native
    rng = np.random.default_rng(4)
    n, d, q, k = 24000, 256, 256, 5
    db = _rows(rng, n, d)
    base = rng.standard_normal((1, d)).astype(np.float32)
    db[:200] = base + 1e-4 * rng.standard_normal((200, d)).astype(np.float32)
    db = db[rng.permutation(n)]
    folds = (np.arange(n) * 5 // n).astype(np.uint8)
    pick = np.linspace(0, n - 1, q).astype(np.int64)
    qrows = db[pick].copy()
    qrows[:8] = base + 1e-4 * rng.standard_normal((8, d)).astype(np.float32)
    t_db = torch.from_numpy(db).to(eng.device)
    full = eng.prepare(t_db, None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    lazy = eng.prepare(t_db, None, 1.0, 1.0, native.NF_ROWNORM, "rescore", defer_f32=True)
    qs = eng.prepare(torch.from_numpy(qrows).to(eng.device), None, 1.0, 1.0, native.NF_ROWNORM, "rescore")
    db_fold = eng.to_device(folds, torch.uint8)
    q_fold = db_fold[torch.from_numpy(pick).to(eng.device)]
    for kw in ({}, {"q_fold": q_fold, "db_fold": db_fold}, {"q_fold": q_fold, "db_fold": db_fold, "fold_sorted": True}):
        want = eng.topk_search(qs, full, k, "rescore", **kw)
        n_full, _ = eng.consume_status()
        got = eng.topk_search(qs, lazy, k, "rescore", **kw)
        n_lazy, _ = eng.consume_status()
        assert n_full == n_lazy and n_full >= 8
        assert torch.equal(got, want)


def test_shapes_the_deferred_arithmetic_does_not_take_fall_back(eng):
    import torch
    from emr2a_b200 import native
    rng = np.random.default_rng(6)
    a = torch.from_numpy(_rows(rng, 3000, 6)).to(eng.device)       # 6 columns: a 4-element chunk would straddle the segments
    b = torch.from_numpy(_rows(rng, 3000, 10)).to(eng.device)
    op = eng.prepare(a, b, 1.0, 1.0, native.NF_SEGNORM | native.NF_ROWNORM, "rescore", defer_f32=True)
    assert op.f32 is not None and op.lazy is None
    with pytest.raises(ValueError):
        eng.normalize_fuse(a, b, 1.0, 1.0, native.NF_ROWNORM, want_planes=True, want_stats=True, defer_f32=True)


def test_search_and_vote_uses_deferred_rows_and_matches_materialised(eng, monkeypatch):
    """The public pipeline call: same keys, votes and counters whether the database rows are deferred or not."""
    import torch
    import emr2a_b200.engine as E
    from emr2a_b200 import native
    rng = np.random.default_rng(12)
    n, d, q, k, c = 40000, 256, 600, 10, 3
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    di, dt = _rows(rng, n, d), _rows(rng, n, d)
    qi, qt = _rows(rng, q, d), _rows(rng, q, d)
    lab, ql = rng.integers(0, c, n).astype(np.int32), rng.integers(0, c, q).astype(np.int32)
    seen = []
    real = eng.normalize_fuse

    def spy(*a, **kw):
        seen.append(bool(kw.get("defer_f32")))
        return real(*a, **kw)
    monkeypatch.setattr(eng, "normalize_fuse", spy)
    r1 = eng.search_and_vote((di, dt), (qi, qt), lab, ql, c, k, db_flags=flags, q_flags=flags, precision="rescore")
    assert seen == [True, False]                                   # database deferred, queries materialised
    monkeypatch.setattr(E, "_DEFER_F32", False)
    seen.clear()
    r0 = eng.search_and_vote((di, dt), (qi, qt), lab, ql, c, k, db_flags=flags, q_flags=flags, precision="rescore")
    assert seen == [False, False]
    for name in ("keys", "pred_vote", "pred_weighted", "hit_counts", "confusion", "top_scores"):
        assert torch.equal(r1[name], r0[name]), name
