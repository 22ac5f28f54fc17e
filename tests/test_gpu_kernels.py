"""Kernel-level parity: every libemr2a.so entry point against the CPU oracle / the golden
vectors produced by the reference.  Needs a B200 (``-m gpu``); all calls go through the
C-ABI (ctypes) via emr2a_b200.engine."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-5          # north_star: fp32-accumulated cosine scores within 1e-5
F32_TOL = 2e-6            # fp32 arms repeat the reference arithmetic; only summation order differs


@pytest.fixture(scope="module")
def eng():
    from emr2a_b200.engine import get_engine
    return get_engine()


def _np(t):
    return t.detach().cpu().numpy()


def _bf16_to_f32(t):
    return (_np(t).view(np.uint16).astype(np.uint32) << 16).view(np.float32)


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("n,d0,d1", [(33, 24, 40), (1000, 512, 512), (7, 25, 0), (64, 1024, 0), (19, 4096, 1024),
                                     (5, 3, 5), (300, 96, 160)])
def test_normalize_fuse_modes(eng, oracle, n, d0, d1):
    from emr2a_b200 import native
    rng = np.random.default_rng(n + d0)
    a = (rng.standard_normal((n, d0)) * rng.uniform(0.1, 5, (n, 1))).astype(np.float32)
    b = (rng.standard_normal((n, d1)) * 2).astype(np.float32) if d1 else None
    if n > 4:
        a[3] = 0
        if b is not None:
            b[3] = 0                                                   # all-zero row: epsilon path
    if b is None:
        got = eng.normalize_fuse(a, flags=native.NF_ROWNORM, want_planes=True, want_inv_norm=True)
        ref = oracle.unit_rows(a)
    else:
        got = eng.normalize_fuse(a, b, flags=native.NF_ROWNORM, want_planes=True, want_inv_norm=True)
        ref = oracle.unit_rows(np.concatenate([a, b], axis=1))
    np.testing.assert_allclose(_np(got.f32), ref, atol=F32_TOL, rtol=0)
    # bf16 hi+lo planes reconstruct the fp32 rows to ~2^-17 relative, pad columns are zero
    d = d0 + d1
    hi, lo = _bf16_to_f32(got.hi), _bf16_to_f32(got.lo)
    assert hi.shape[1] % 64 == 0 and np.all(hi[:, d:] == 0) and np.all(lo[:, d:] == 0)
    np.testing.assert_allclose(hi[:, :d] + lo[:, :d], _np(got.f32), atol=1e-6 * max(1e-3, float(np.abs(ref).max())) * 16, rtol=2e-5)
    if b is not None:
        # weighted early fusion (text first) and the CV chain normalise-each -> concat -> normalise
        w = eng.normalize_fuse(a, b, np.float32(0.4), np.float32(0.6), native.NF_ROWNORM)
        np.testing.assert_allclose(_np(w.f32), oracle.fuse_early(a, b, 0.4, 0.6), atol=F32_TOL, rtol=0)
        c = eng.normalize_fuse(a, b, flags=native.NF_SEGNORM | native.NF_ROWNORM)
        np.testing.assert_allclose(_np(c.f32), oracle.fuse_concat_cv(oracle.unit_rows(a), oracle.unit_rows(b)),
                                   atol=F32_TOL, rtol=0)


def test_normalize_fuse_bf16_input_and_strides(eng, oracle):
    import torch
    from emr2a_b200 import native
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((50, 128)).astype(np.float32)).cuda().to(torch.bfloat16)
    got = eng.normalize_fuse(x, flags=native.NF_ROWNORM)
    ref = oracle.unit_rows(x.float().cpu().numpy())
    np.testing.assert_allclose(_np(got.f32), ref, atol=F32_TOL, rtol=0)
    # zero guard (utils/common.py): zero vector stays zero, no epsilon
    z = eng.normalize_fuse(np.zeros((1, 7), np.float32), flags=native.NF_ZERO_GUARD)
    assert np.all(_np(z.f32) == 0)
    v = rng.standard_normal((1, 7)).astype(np.float32)
    np.testing.assert_allclose(_np(eng.normalize_fuse(v, flags=native.NF_ZERO_GUARD).f32)[0], oracle.unit_vector(v[0]),
                               atol=F32_TOL)


def test_golden_primitives_through_python_api(golden):
    """The reference's primitive functions, same names/signatures, against reference outputs."""
    from emr2a_b200.retrieval import compute_cosine_similarity, compute_euclidean_similarity, late_fusion, early_fusion
    from emr2a_b200.retrieval.fusion import normalize_scores
    from emr2a_b200.utils import l2_normalize, concat_embeddings
    from emr2a_b200.utils.cv_evaluator import CVRetrievalEvaluator
    g = golden("primitives.npz")
    np.testing.assert_allclose(compute_cosine_similarity(g["cos_q"], g["cos_db"]), g["cos_out"], atol=F32_TOL)
    np.testing.assert_allclose(compute_euclidean_similarity(g["cos_q"], g["cos_db"]), g["euc_out"], atol=F32_TOL)
    np.testing.assert_allclose(early_fusion(g["ef_text"], g["ef_image"]), g["ef_out_11"], atol=F32_TOL)
    np.testing.assert_allclose(early_fusion(g["ef_text"], g["ef_image"], 0.4, 0.6), g["ef_out_w"], atol=F32_TOL)
    for mode in ("none", "zscore", "minmax"):
        np.testing.assert_allclose(late_fusion(g["lf_ts"], g["lf_is"], 0.4, mode), g[f"lf_out_{mode}"], atol=1e-5, rtol=1e-5)
        np.testing.assert_allclose(normalize_scores(g["lf_ts"], mode), g[f"ns_out_{mode}"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(late_fusion(g["lf_ts"], g["lf_is"], 0.7), g["lf_out_w07"], atol=F32_TOL)
    ev = CVRetrievalEvaluator()
    np.testing.assert_allclose(ev._normalize_rows(g["ef_image"]), g["nr_out"], atol=F32_TOL)
    np.testing.assert_allclose(ev.concat_fusion(g["cf_img"], g["cf_txt"]), g["cf_out"], atol=F32_TOL)
    np.testing.assert_allclose(ev.compute_cosine_similarity(g["cf_out"][2], g["cf_out"]), g["dot_out"], atol=F32_TOL)
    np.testing.assert_allclose(l2_normalize(g["cos_q"]), g["l2_out"], atol=F32_TOL)
    assert np.array_equal(l2_normalize(np.zeros(7, np.float32)), g["l2_zero_out"])
    np.testing.assert_allclose(concat_embeddings(g["ef_text"][0], g["ef_image"][0], 0.3, 0.9), g["ce_out"], atol=F32_TOL)
    labels = [f"class_{c}" for c in g["rt_labels"]]
    tl, ts, tid = ev.retrieve_topk(g["cf_out"][2], g["cf_out"], labels, 5)
    assert [int(x.split("_")[1]) for x in tid] == list(g["rt_top_idx"])
    assert [int(x.split("_")[1]) for x in tl] == list(g["rt_top_labels"])
    np.testing.assert_allclose(ts, g["rt_top_scores"], atol=F32_TOL)
    # votes on given lists
    vl = [[f"class_{c}" for c in r] for r in g["v_labels"]]
    vs = [[float(x) for x in r] for r in g["v_scores"]]
    true = [f"class_{c}" for c in g["v_true"]]
    assert ev.compute_vote_accuracy(vl, vs, true, weighted=False) == float(g["v_acc_major"])
    assert ev.compute_vote_accuracy(vl, vs, true, weighted=True) == float(g["v_acc_weight"])


# ------------------------------------------------------------------ K2 (fp32 arm)
def _check_topk(oracle, keys, qs, db, k, q_fold=None, db_fold=None, tol=SCORE_TOL, idx_base=0, min_safe=0.5):
    from emr2a_b200.engine import unpack_keys
    sc, idx = unpack_keys(keys)
    o_idx, o_sc = oracle.search_topk_batched(qs, db, k, q_fold=q_fold, db_fold=db_fold)
    valid = o_idx >= 0
    assert np.array_equal(idx >= 0, valid)
    assert np.max(np.abs(np.where(valid, sc - o_sc, 0))) <= tol
    # index sets bit-exact wherever adjacent gaps (including the gap to the first loser) exceed 2*tol
    full = qs.astype(np.float64) @ db.astype(np.float64).T
    if q_fold is not None:
        full = np.where(q_fold[:, None] == db_fold[None, :], -np.inf, full)
    srt = -np.sort(-full, axis=1)[:, :k + 1]
    gaps = np.abs(np.diff(srt, axis=1))
    gaps = np.where(np.isfinite(gaps), gaps, np.inf)
    safe = gaps.min(axis=1) > 2 * tol if gaps.shape[1] else np.ones(len(qs), bool)
    assert safe.mean() >= min_safe
    assert np.array_equal(np.where(valid, idx - idx_base, -1)[safe], o_idx[safe])
    return safe


@pytest.mark.parametrize("Q,N,D,K", [(1, 1, 8, 1), (70, 300, 33, 5), (130, 1000, 256, 10), (64, 129, 64, 40),
                                     (5, 3, 16, 5), (257, 4097, 100, 3)])
def test_topk_search_fp32(eng, oracle, Q, N, D, K):
    rng = np.random.default_rng(Q * 7 + N)
    db = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    qs = oracle.unit_rows(rng.standard_normal((Q, D)).astype(np.float32))
    if N > 20:
        db[7] = db[11]                                   # exact duplicate rows: tie -> lower index first
        qs[0] = db[7]
    keys = eng.topk_search(eng.prepare(qs, flags=0), eng.prepare(db, flags=0), K, "fp32")
    _check_topk(oracle, keys, qs, db, K)
    if N > 20:
        from emr2a_b200.engine import unpack_keys
        _, idx = unpack_keys(keys)
        assert list(idx[0][:2]) == [7, 11]


def test_topk_search_fp32_fold_mask_and_base(eng, oracle):
    rng = np.random.default_rng(1)
    N, Q, D, K = 2000, 200, 64, 5
    db = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    fold = rng.integers(0, 5, N).astype(np.uint8)
    qs, qf = db[:Q].copy(), fold[:Q].copy()            # queries ARE database rows, as in the CV loop
    import torch
    keys = eng.topk_search(eng.prepare(qs, flags=0), eng.prepare(db, flags=0), K, "fp32",
                           q_fold=torch.from_numpy(qf), db_fold=torch.from_numpy(fold), idx_base=1000)
    _check_topk(oracle, keys, qs, db, K, q_fold=qf, db_fold=fold, idx_base=1000)


# ------------------------------------------------------------------ K2 (tcgen05 arms)
@pytest.mark.parametrize("prec,tol", [("bf16x3", SCORE_TOL), ("rescore", F32_TOL)])
@pytest.mark.parametrize("Q,N,D,K", [(128, 256, 64, 5), (100, 1000, 128, 5), (300, 5000, 200, 10),
                                     (257, 3333, 96, 20), (1000, 40000, 1024, 10), (2, 70, 8, 3), (40, 20, 100, 16)])
def test_topk_search_tensor_core(eng, oracle, prec, tol, Q, N, D, K):
    if prec == "rescore" and K > 10:
        pytest.skip("rescore arm serves K <= 10")
    rng = np.random.default_rng(Q + N + D)
    db = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    qs = oracle.unit_rows(rng.standard_normal((Q, D)).astype(np.float32))
    if N >= 64:
        db[5] = db[9]
        qs[1] = db[5]
    keys = eng.topk_search(eng.prepare(qs, flags=0, precision=prec), eng.prepare(db, flags=0, precision=prec), K, prec)
    _check_topk(oracle, keys, qs, db, K, tol=tol)
    if N >= 64:
        from emr2a_b200.engine import unpack_keys
        _, idx = unpack_keys(keys)
        assert list(idx[1][:2]) == [5, 9]               # identical rows give identical scores: index order decides


@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "rescore"])
def test_duplicate_rows_resolve_to_the_lowest_index(eng, oracle, prec):
    """The documented deviation from the reference (include/emr2a.h "packed key"): among EQUAL
    scores the lowest database index wins, on every arm and across the K boundary.  The reference's
    `np.argsort(x)[-k:][::-1]` (utils/cv_evaluator.py:123) leaves that order to numpy's sort -- with a stable sort it
    lists the HIGHEST index first -- so for duplicated embeddings (identical reports in text_only fusion) its Top-K
    membership at the boundary is one of several equally valid answers; ours is fixed and independent of tile shape,
    split count and GPU count."""
    from emr2a_b200.engine import unpack_keys
    rng = np.random.default_rng(77)
    N, D, K = 50000, 128, 3
    db = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    dups = [10, 500, 900, 40000, 49999]                     # five copies of one case, spread over tiles and splits
    for j in dups[1:]:
        db[j] = db[dups[0]]
    qs = oracle.unit_rows(rng.standard_normal((300, D)).astype(np.float32))
    qs[7] = db[dups[0]]
    keys = eng.topk_search(eng.prepare(qs, flags=0, precision=prec), eng.prepare(db, flags=0, precision=prec), K, prec)
    if prec == "rescore":
        eng.consume_status()
    scores, idx = unpack_keys(keys)
    assert list(idx[7]) == dups[:K]                        # K = 3 of the 5 tied rows: the three LOWEST indices
    assert scores[7][0] == scores[7][1] == scores[7][2]
    # what numpy does with the same scores: also three of the five copies -- but which three is the sort's business
    exact = (db.astype(np.float64) @ qs[7].astype(np.float64)).astype(np.float32)
    ref = np.argsort(exact)[-K:][::-1]
    assert set(ref) <= set(dups)
    stable = np.argsort(exact, kind="stable")[-K:][::-1]
    assert list(stable) == dups[::-1][:K]                   # a stable sort would have kept the three HIGHEST


def test_rescore_unverifiable_queries_are_rescanned_exactly(eng, oracle):
    """Adversarial database: for some queries ~100 rows sit within 5e-3 of the best score (5e-5 apart),
    far inside the bf16 filter's error bound, so the bound cannot verify the selection -> those
    queries go through the exact re-scan and must still match the oracle index for index."""
    import torch
    rng = np.random.default_rng(99)
    N, Q, D, K = 20000, 64, 256, 10
    db = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    qs = oracle.unit_rows(rng.standard_normal((Q, D)).astype(np.float32))
    hard = [3, 17, 40]
    for h, qi in enumerate(hard):
        for j in range(100):
            u = rng.standard_normal(D)
            u -= u.dot(qs[qi]) * qs[qi]
            u /= np.linalg.norm(u)
            a = np.sqrt(2 * 5e-5 * (j + 1))
            v = qs[qi].astype(np.float64) + a * u
            db[1000 * (h + 1) + 7 * j] = (v / np.linalg.norm(v)).astype(np.float32)
    fold = rng.integers(0, 5, N).astype(np.uint8)
    qf = rng.integers(0, 5, Q).astype(np.uint8)
    for use_fold in (False, True):
        kw = dict(q_fold=torch.from_numpy(qf), db_fold=torch.from_numpy(fold)) if use_fold else {}
        keys = eng.topk_search(eng.prepare(qs, flags=0, precision="rescore"), eng.prepare(db, flags=0, precision="rescore"),
                               K, "rescore", **kw)
        unverified, overflow = eng.consume_status()
        assert unverified >= len(hard) and not overflow
        _check_topk(oracle, keys, qs, db, K, tol=F32_TOL, **({"q_fold": qf, "db_fold": fold} if use_fold else {}))


def test_rescore_overflow_falls_back_to_three_pass(eng, oracle):
    """Every query unverifiable (database = one tight cluster) and more queries than the re-scan list
    holds: the engine must notice the overflow flag and re-search the unverified queries with the BF16X3 arm."""
    from emr2a_b200.engine import unpack_keys
    rng = np.random.default_rng(5)
    N, Q, D, K = 3000, 1500, 128, 5
    c = rng.standard_normal(D)
    db = oracle.unit_rows((c + 0.05 * rng.standard_normal((N, D))).astype(np.float32))
    qs = oracle.unit_rows((c + 0.05 * rng.standard_normal((Q, D))).astype(np.float32))
    labels = rng.integers(0, 3, N).astype(np.int32)
    r = eng.search_and_vote((db,), (qs,), labels, labels[:Q], 3, K, db_flags=0, q_flags=0, precision="rescore")
    assert r["precision"] == "rescore+bf16x3" and r["unverified"] > 1400
    _check_topk(oracle, r["keys"], qs, db, K, tol=SCORE_TOL, min_safe=0.0)     # a tight cluster has few clear gaps
    # below the capacity (64 for small batches) the exact re-scan handles all of them
    r = eng.search_and_vote((db,), (qs[:60],), labels, labels[:60], 3, K, db_flags=0, q_flags=0, precision="rescore")
    assert r["precision"] == "rescore" and r["unverified"] > 40
    _check_topk(oracle, r["keys"], qs[:60], db, K, tol=F32_TOL, min_safe=0.0)


def test_topk_search_bf16x1_matches_bf16_math(eng, oracle):
    """bf16-input variant: products of bf16 values are exact, accumulation is fp32 -- compare with the
    same bf16-rounded operands evaluated in float64 (tolerance 1e-5), NOT with the unrounded fp32 inputs."""
    rng = np.random.default_rng(17)
    Q, N, D, K = 200, 6000, 320, 5
    db = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    qs = oracle.unit_rows(rng.standard_normal((Q, D)).astype(np.float32))
    qo, do = eng.prepare(qs, flags=0, precision="bf16x1"), eng.prepare(db, flags=0, precision="bf16x1")
    keys = eng.topk_search(qo, do, K, "bf16x1")
    qh, dh = _bf16_to_f32(qo.hi)[:, :D], _bf16_to_f32(do.hi)[:, :D]
    _check_topk(oracle, keys, qh, dh, K, tol=SCORE_TOL)


def test_topk_search_tensor_core_fold_mask(eng, oracle):
    import torch
    rng = np.random.default_rng(23)
    N, Q, D, K = 6000, 300, 192, 5
    db = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    fold = np.sort(rng.integers(0, 5, N)).astype(np.uint8)
    pick = rng.choice(N, Q, replace=False)
    qs, qf = db[pick].copy(), fold[pick].copy()
    keys = eng.topk_search(eng.prepare(qs, flags=0, precision="bf16x3"), eng.prepare(db, flags=0, precision="bf16x3"),
                           K, "bf16x3", q_fold=torch.from_numpy(qf), db_fold=torch.from_numpy(fold))
    _check_topk(oracle, keys, qs, db, K, q_fold=qf, db_fold=fold)


def test_tensor_core_equals_fp32_arm_large(eng):
    """Property at a size the CPU oracle would take minutes for: both arms agree on the index
    sets (where the fp32 arm's gaps are clear) and on scores to 1e-5."""
    import torch
    from emr2a_b200 import synth, native
    from emr2a_b200.engine import unpack_keys
    data = synth.two_modal(200_000 + 2048, 256, 256, 3, seed=11)
    db = (data["image"][:200_000], data["text"][:200_000])
    qs = (data["image"][200_000:], data["text"][200_000:])
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    res = {}
    for prec in ("fp32", "bf16x3", "rescore"):
        res[prec] = unpack_keys(eng.topk_search(eng.prepare(qs[0], qs[1], flags=flags, precision=prec),
                                                eng.prepare(db[0], db[1], flags=flags, precision=prec), 10, prec))
    (s32, i32), (s3, i3), (sr, ir) = res["fp32"], res["bf16x3"], res["rescore"]
    assert np.max(np.abs(s32 - s3)) < SCORE_TOL
    safe = np.abs(np.diff(s32, axis=1)).min(axis=1) > 2 * SCORE_TOL
    assert safe.mean() > 0.8
    assert np.array_equal(i32[safe][:, :9], i3[safe][:, :9])
    # the rescore arm re-computes the scores in fp32 from the same rows as the fp32 arm
    assert np.max(np.abs(s32 - sr)) < F32_TOL
    safe = np.abs(np.diff(s32, axis=1)).min(axis=1) > 2 * F32_TOL
    assert np.array_equal(i32[safe][:, :9], ir[safe][:, :9])
    assert eng.consume_status() == (0, False)


# ------------------------------------------------------------------ K3 / K4
def test_topk_merge(eng):
    import torch
    rng = np.random.default_rng(2)
    for parts, Q, K_in, K_out in [(2, 100, 5, 5), (8, 37, 10, 10), (13, 50, 16, 7), (40, 9, 32, 32), (3, 5, 4, 9)]:
        raw = rng.integers(1, 2**62, size=(parts, Q, K_in), dtype=np.int64)
        raw = -np.sort(-raw, axis=2)
        raw[0, 0, K_in - 1] = 0                              # an empty slot
        got = _np(eng.topk_merge(torch.from_numpy(raw).cuda(), K_out))
        flat = np.transpose(raw, (1, 0, 2)).reshape(Q, -1)
        want = -np.sort(-flat, axis=1)[:, :K_out]
        if K_out > flat.shape[1]:
            want = np.concatenate([want, np.zeros((Q, K_out - flat.shape[1]), np.int64)], axis=1)
        assert np.array_equal(got, want)


def test_vote_metrics_against_oracle(eng, oracle):
    import torch
    rng = np.random.default_rng(4)
    Q, K, C, N = 500, 5, 3, 1000
    db_labels = rng.integers(0, C, N).astype(np.int32)
    q_labels = rng.integers(0, C, Q).astype(np.int32)
    groups = rng.integers(0, 5, Q).astype(np.uint8)
    idx = np.stack([rng.choice(N, K, replace=False) for _ in range(Q)])
    sc = -np.sort(-rng.random((Q, K)).astype(np.float32), axis=1)
    sc[::9, 1] = sc[::9, 0]                                   # equal scores
    sc[::11, 2:] = sc[::11, 1:2]                              # equal sums are likely -> first-label rule
    bits = sc.view(np.uint32).astype(np.uint64) ^ np.uint64(0x80000000)
    keys = (bits << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - idx.astype(np.uint64))
    for wacc in (False, True):
        r = eng.vote_metrics(torch.from_numpy(keys.view(np.int64)).cuda(), db_labels, q_labels, C, k_list=[1, 3, 5, 5],
                             wacc_f32=wacc, q_group=torch.from_numpy(groups), n_groups=5)
        labs = db_labels[idx]
        maj = np.array([oracle.vote_majority([int(x) for x in row]) for row in labs])
        wv = np.array([oracle.vote_weighted([int(x) for x in row], list(s), "f32" if wacc else "f64")
                       for row, s in zip(labs, sc)])
        assert np.array_equal(_np(r["pred_top1"]), labs[:, 0])
        assert np.array_equal(_np(r["pred_vote"]), maj)
        assert np.array_equal(_np(r["pred_weighted"]), wv)
        assert np.array_equal(_np(r["top_idx"]), idx) and np.array_equal(_np(r["top_labels"]), labs)
        assert np.array_equal(_np(r["top_scores"]), sc)
        for g in range(5):
            m = groups == g
            assert int(r["group_sizes"][g]) == int(m.sum())
            for j, k in enumerate([1, 3, 5, 5]):
                assert int(r["hit_counts"][g, j]) == int(sum(q_labels[i] in labs[i, :k] for i in np.nonzero(m)[0]))
            assert int(r["vote_counts"][g, 1]) == int((maj[m] == q_labels[m]).sum())
            assert int(r["vote_counts"][g, 2]) == int((wv[m] == q_labels[m]).sum())
            assert np.array_equal(_np(r["confusion"][g, 0]), oracle.confusion_counts(labs[m, 0], q_labels[m], C))
            assert np.array_equal(_np(r["confusion"][g, 1]), oracle.confusion_counts(maj[m], q_labels[m], C))


def test_topk_from_scores(eng, oracle):
    rng = np.random.default_rng(8)
    s = rng.standard_normal((40, 333)).astype(np.float32)
    s[:, 10] = s[:, 20]
    from emr2a_b200.engine import unpack_keys
    sc, idx = unpack_keys(eng.topk_from_scores(s, 7))
    for i in range(40):
        assert np.array_equal(idx[i], oracle.topk_desc(s[i], 7))
    sc, idx = unpack_keys(eng.topk_from_scores(s[:, :3], 5))         # K > N
    assert np.all(idx[:, 3:] == -1) and np.all(idx[:, :3] >= 0)


def test_errors_are_loud(eng):
    from emr2a_b200 import native
    with pytest.raises(ValueError):
        eng.normalize_fuse(np.zeros((3, 4), np.float32), np.zeros((2, 4), np.float32))
    with pytest.raises(ValueError):
        eng.topk_search(eng.prepare(np.ones((2, 8), np.float32)), eng.prepare(np.ones((2, 16), np.float32)), 1)
    rc = eng.lib.emr2a_scores(None, None, 1, 1, 4, 4, 4, None, 1, None)
    assert rc == native.ERR_INVALID and "scores" in native.last_error()


# ------------------------------------------------------------------ CV rule at scale / other BASELINE configs
@pytest.mark.parametrize("prec", ["bf16x3", "rescore", "fp32"])
def test_fold_sorted_tile_skipping(eng, oracle, prec):
    """Every row a query, rows in fold order, fold_sorted=1: whole 128x256 tiles of the query's own fold
    are skipped in the tcgen05 kernel -- results must equal the element-masked search."""
    import torch
    rng = np.random.default_rng(31)
    N, D, K = 6400, 128, 5
    db = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    fold = np.sort(rng.integers(0, 5, N)).astype(np.uint8)
    op = eng.prepare(db, flags=0, precision=prec)
    f = torch.from_numpy(fold)
    keys = eng.topk_search(op, op, K, prec, q_fold=f, db_fold=f, fold_sorted=True)
    eng.consume_status()
    _check_topk(oracle, keys, db, db, K, q_fold=fold, db_fold=fold, tol=SCORE_TOL if prec == "bf16x3" else F32_TOL)
    keys2 = eng.topk_search(op, op, K, prec, q_fold=f, db_fold=f, fold_sorted=False)
    eng.consume_status()
    assert torch.equal(keys, keys2)


@pytest.mark.parametrize("prec", ["rescore", "bf16x3"])
def test_cv_all_folds_engine_path(eng, oracle, prec):
    """Engine.cv_search_and_vote with UNSORTED folds: rows are permuted into fold order internally and
    everything is mapped back; compare with the oracle's masked search + votes, per fold counters."""
    rng = np.random.default_rng(77)
    N, D, K, C = 5000, 96, 5, 3
    img = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    txt = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    labels = rng.integers(0, C, N).astype(np.int32)
    fold = rng.integers(0, 5, N).astype(np.uint8)
    from emr2a_b200 import native
    r = eng.cv_search_and_vote((img, txt), labels, fold, C, K, flags=native.NF_ROWNORM, k_list=[1, 3, 5],
                               precision=prec, n_folds=5, q_block=2048)
    fused = oracle.fuse_concat_cv(img, txt)
    o_idx, o_sc = oracle.search_topk_batched(fused, fused, K, q_fold=fold, db_fold=fold)
    got_idx, got_sc = _np(r["top_idx"]), _np(r["top_scores"])
    tol = SCORE_TOL if prec == "bf16x3" else F32_TOL
    assert np.max(np.abs(got_sc - o_sc)) < tol
    safe = np.abs(np.diff(o_sc, axis=1)).min(axis=1) > 2 * tol
    assert safe.mean() > 0.9 and np.array_equal(got_idx[safe], o_idx[safe])
    assert np.all(fold[got_idx] != fold[:, None])                       # never retrieved from the own fold
    maj = np.array([oracle.vote_majority([int(x) for x in labels[row]]) for row in o_idx])
    assert np.array_equal(_np(r["pred_vote"])[safe], maj[safe])
    for f in range(5):
        m = fold == f
        assert int(r["group_sizes"][f]) == int(m.sum())
        if safe[m].all():
            assert int(r["vote_counts"][f, 1]) == int((maj[m] == labels[m]).sum())
            assert int(r["hit_counts"][f, 0]) == int((labels[o_idx[m, 0]] == labels[m]).sum())


def test_c3_late_fusion_merge_then_topk(eng, oracle):
    """BASELINE config C3 shape (512-d image + 512-d text, late fusion): the reference merges the two
    score vectors of ALL rows and then ranks (utils/cv_evaluator.py:233-237); here the weights are
    folded into the query rows and one contraction does both."""
    rng = np.random.default_rng(13)
    N, Q, D, K = 50_000, 512, 512, 10
    for w in (0.25, 0.5):
        di = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
        dt = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
        qi = oracle.unit_rows(rng.standard_normal((Q, D)).astype(np.float32))
        qt = oracle.unit_rows(rng.standard_normal((Q, D)).astype(np.float32))
        ref = w * (qt.astype(np.float64) @ dt.T.astype(np.float64)) + (1 - w) * (qi.astype(np.float64) @ di.T.astype(np.float64))
        o_idx = np.argsort(-ref, axis=1, kind="stable")[:, :K]
        o_sc = np.take_along_axis(ref, o_idx, axis=1)
        from emr2a_b200.engine import unpack_keys
        for prec in ("rescore", "bf16x3"):
            db = eng.prepare(di, dt, 1.0, 1.0, 0, prec)
            qs = eng.prepare(qi, qt, np.float32(1 - w), np.float32(w), 0, prec)
            sc, idx = unpack_keys(eng.topk_search(qs, db, K, prec))
            assert eng.consume_status() == (0, False)
            tol = SCORE_TOL if prec == "bf16x3" else F32_TOL
            assert np.max(np.abs(sc - o_sc)) < tol
            safe = np.abs(np.diff(np.sort(-ref, axis=1)[:, :K + 1], axis=1)).min(axis=1) > 2 * tol
            assert safe.mean() > 0.7 and np.array_equal(idx[safe], o_idx[safe])


def test_c4_bf16_inputs_wide_rows(eng, oracle):
    """BASELINE config C4 shape: 4096-d image + 1024-d text, bf16 INPUTS, fp32 accumulation.  The oracle
    consumes the bf16 values up-cast to fp32."""
    import torch
    from emr2a_b200 import native
    from emr2a_b200.engine import unpack_keys
    rng = np.random.default_rng(17)
    N, Q, K = 12_000, 200, 5
    to_bf16 = lambda a: torch.from_numpy(a).to(torch.bfloat16)                      # noqa: E731
    di, dt = to_bf16(rng.standard_normal((N, 4096)).astype(np.float32)), to_bf16(rng.standard_normal((N, 1024)).astype(np.float32))
    qi, qt = to_bf16(rng.standard_normal((Q, 4096)).astype(np.float32)), to_bf16(rng.standard_normal((Q, 1024)).astype(np.float32))
    up = lambda t: t.float().numpy()                                                 # noqa: E731
    odb = oracle.fuse_concat_cv(oracle.unit_rows(up(di)), oracle.unit_rows(up(dt)))
    oq = oracle.fuse_concat_cv(oracle.unit_rows(up(qi)), oracle.unit_rows(up(qt)))
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    for prec, tol in (("rescore", F32_TOL), ("bf16x3", SCORE_TOL), ("fp32", F32_TOL)):
        db = eng.prepare(di.cuda(), dt.cuda(), 1.0, 1.0, flags, prec)
        qs = eng.prepare(qi.cuda(), qt.cuda(), 1.0, 1.0, flags, prec)
        keys = eng.topk_search(qs, db, K, prec)
        eng.consume_status()
        _check_topk(oracle, keys, oq, odb, K, tol=tol)


def test_ingest_mean_pool_and_loaders(eng, tmp_path):
    """Slice mean-pool on the GPU equals numpy's ``arr.mean(axis=0)`` per patient bit for bit (the ingest of
    pipelines/step3_retrieval/evaluate_retrieval.py:66-67), for ragged slice counts; both .npz layouts load."""
    from emr2a_b200 import ingest
    rng = np.random.default_rng(3)
    per = [rng.standard_normal((int(rng.integers(1, 9)), 96)).astype(np.float32) * 3 for _ in range(57)]
    per.append(rng.standard_normal(96).astype(np.float32))            # already pooled (ndim 1)
    got = _np(ingest.mean_pool_patients(per))
    want = np.stack([np.atleast_2d(a).mean(axis=0) for a in per])
    assert got.dtype == np.float32 and np.array_equal(got, want)
    ids = [f"p{i:03d}" for i in range(len(per) - 1)]
    np.savez(tmp_path / "step2.npz", **{pid: a for pid, a in zip(ids, per[:-1])})
    ids2, pooled = ingest.load_patient_npz(tmp_path / "step2.npz")
    assert ids2 == ids and np.array_equal(_np(pooled), want[:-1])
    np.savez(tmp_path / "cv.npz", patient_ids=np.array(ids, dtype=object), image_matrix=want[:-1], text_matrix=want[:-1] * 2)
    m = ingest.load_matrix_npz(tmp_path / "cv.npz")
    assert m["patient_ids"] == ids and np.array_equal(m["image"], want[:-1]) and np.array_equal(m["text"], want[:-1] * 2)
    emb = {pid: {"image": want[i], "text": want[i] * 2} for i, pid in enumerate(ids)}
    a, b = ingest.embeddings_to_arrays(ids, emb)
    assert np.array_equal(a, want[:-1]) and np.array_equal(b, want[:-1] * 2)


def test_rescore_on_clustered_database_matches_fp32_arm(eng):
    """Real embedding sets are clustered (near-duplicate cases, often stored contiguously).  Small clusters fit
    into the per-split candidate lists and verify.  Clusters larger than a split's list (16/32 rows within the
    filter's error band 2E ~ 8e-3) cannot be verified: those queries are flagged, the re-scan list overflows
    and the engine re-searches the flagged queries with the BF16X3 arm.  Either way the result must equal the
    fp32 arm's."""
    import torch
    from emr2a_b200 import native
    from emr2a_b200.engine import unpack_keys
    g = torch.Generator(device="cuda").manual_seed(5)
    D, Q, K = 256, 1500, 10
    for n_clusters, per, noise, expect in ((2500, 8, 0.05, "rescore"), (400, 250, 0.02, "rescore+bf16x3")):
        centres = torch.randn((n_clusters, D), generator=g, device="cuda")
        db = centres.repeat_interleave(per, dim=0) + noise * torch.randn((n_clusters * per, D), generator=g, device="cuda")
        pick = torch.randint(0, n_clusters, (Q,), generator=g, device="cuda")
        qs = centres[pick] + noise * torch.randn((Q, D), generator=g, device="cuda")
        labels = torch.zeros((n_clusters * per,), dtype=torch.int32, device="cuda")
        ql = torch.zeros((Q,), dtype=torch.int32, device="cuda")
        ref = eng.search_and_vote((db,), (qs,), labels, ql, 1, K, precision="fp32")
        got = eng.search_and_vote((db,), (qs,), labels, ql, 1, K, precision="rescore")
        assert got["precision"] == expect
        (s32, i32), (sr, ir) = unpack_keys(ref["keys"]), unpack_keys(got["keys"])
        tol = F32_TOL if expect == "rescore" else SCORE_TOL
        assert np.max(np.abs(s32 - sr)) < tol
        assert np.array_equal(i32 // per, ir // per)             # neighbours come from the query's own cluster
        gaps = np.abs(np.diff(s32, axis=1)).min(axis=1) > 2 * tol
        assert np.array_equal(i32[gaps], ir[gaps])


def test_rescore_partial_overflow_only_flagged_queries_are_redone(eng):
    """A batch where a minority of the queries sits in large tight clusters: those (and only those) are
    re-searched with BF16X3; everything else keeps its verified fp32 result."""
    import torch
    from emr2a_b200 import native
    from emr2a_b200.engine import unpack_keys
    g = torch.Generator(device="cuda").manual_seed(11)
    D, K = 128, 5
    base = torch.randn((60_000, D), generator=g, device="cuda")
    c = torch.randn((4, D), generator=g, device="cuda")
    blob = c.repeat_interleave(300, dim=0) + 0.01 * torch.randn((1200, D), generator=g, device="cuda")
    db = torch.cat([base, blob])
    q_easy = torch.randn((1400, D), generator=g, device="cuda")
    q_hard = c[torch.randint(0, 4, (200,), generator=g, device="cuda")] + 0.01 * torch.randn((200, D), generator=g, device="cuda")
    qs = torch.cat([q_easy, q_hard])
    labels = torch.zeros((db.shape[0],), dtype=torch.int32, device="cuda")
    ql = torch.zeros((qs.shape[0],), dtype=torch.int32, device="cuda")
    ref = eng.search_and_vote((db,), (qs,), labels, ql, 1, K, precision="fp32")
    got = eng.search_and_vote((db,), (qs,), labels, ql, 1, K, precision="rescore")
    assert got["precision"] == "rescore+bf16x3" and 200 <= got["unverified"] < 400
    (s32, i32), (sr, ir) = unpack_keys(ref["keys"]), unpack_keys(got["keys"])
    assert np.max(np.abs(s32 - sr)[:1400]) < F32_TOL and np.max(np.abs(s32 - sr)) < SCORE_TOL
    gaps = np.abs(np.diff(s32, axis=1)).min(axis=1) > 2 * SCORE_TOL
    assert np.array_equal(i32[gaps], ir[gaps]) and gaps[:1400].mean() > 0.95
    assert np.all(ir[1400:] >= 60_000)                         # hard queries retrieve from the blobs


def test_rescore_bound_is_per_query(eng):
    """One query 50x longer than the rest of the batch (per-query scaled late-fusion queries look like this): its
    own error bound is 50x wider, the unit queries keep theirs -- no mass re-scan, and every row equals the fp32 arm."""
    from emr2a_b200 import synth
    from emr2a_b200.engine import unpack_keys
    data = synth.two_modal(40_000 + 3000, 96, 96, 3, seed=77)
    db = np.concatenate([data["image"][:40_000], data["text"][:40_000]], axis=1)
    qs = np.concatenate([data["image"][40_000:], data["text"][40_000:]], axis=1)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    qs[1234] *= 50.0
    res = {}
    for prec in ("fp32", "rescore"):
        res[prec] = unpack_keys(eng.topk_search(eng.prepare(qs, flags=0, precision=prec), eng.prepare(db, flags=0, precision=prec),
                                                10, prec))
    unverified, overflow = eng.consume_status()
    assert not overflow and unverified <= 3, (unverified, overflow)
    (s32, i32), (sr, ir) = res["fp32"], res["rescore"]
    assert np.max(np.abs(s32 - sr) / np.maximum(1.0, np.abs(s32))) < F32_TOL
    clear = np.abs(np.diff(s32, axis=1)).min(axis=1) > 2 * F32_TOL * np.maximum(1.0, np.abs(s32[:, 0]))
    assert clear.mean() > 0.9 and np.array_equal(i32[clear], ir[clear])
