"""Pin the CPU oracle against vectors produced by running the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

TOL = 2e-6   # oracle repeats the same numpy ops; only BLAS thread order may differ


def test_primitives(golden, oracle):
    g = golden("primitives.npz")
    np.testing.assert_allclose(oracle.cosine_one_vs_db(g["cos_q"], g["cos_db"]), g["cos_out"], atol=TOL)
    np.testing.assert_allclose(oracle.euclid_one_vs_db(g["cos_q"], g["cos_db"]), g["euc_out"], atol=TOL)
    np.testing.assert_allclose(oracle.fuse_early(g["ef_text"], g["ef_image"]), g["ef_out_11"], atol=TOL)
    np.testing.assert_allclose(oracle.fuse_early(g["ef_text"], g["ef_image"], 0.4, 0.6), g["ef_out_w"], atol=TOL)
    for mode in ("none", "zscore", "minmax"):
        np.testing.assert_allclose(oracle.fuse_late_scores(g["lf_ts"], g["lf_is"], 0.4, mode),
                                   g[f"lf_out_{mode}"], atol=TOL)
        np.testing.assert_allclose(oracle.rescale_scores(g["lf_ts"], mode), g[f"ns_out_{mode}"], atol=TOL)
    np.testing.assert_allclose(oracle.fuse_late_scores(g["lf_ts"], g["lf_is"], 0.7), g["lf_out_w07"], atol=TOL)
    np.testing.assert_allclose(oracle.unit_rows(g["ef_image"]), g["nr_out"], atol=TOL)
    np.testing.assert_allclose(oracle.fuse_concat_cv(g["cf_img"], g["cf_txt"]), g["cf_out"], atol=TOL)
    np.testing.assert_allclose(oracle.dot_one_vs_db(g["cf_out"][2], g["cf_out"]), g["dot_out"], atol=TOL)
    np.testing.assert_allclose(oracle.unit_vector(g["cos_q"]), g["l2_out"], atol=TOL)
    assert np.array_equal(oracle.unit_vector(np.zeros(7, np.float32)), g["l2_zero_out"])
    np.testing.assert_allclose(oracle.fuse_single(g["ef_text"][0], g["ef_image"][0], 0.3, 0.9), g["ce_out"], atol=TOL)
    assert g["cos_out"].dtype == np.float32 and g["ef_out_11"].dtype == np.float32


def test_retrieve_topk(golden, oracle):
    g = golden("primitives.npz")
    sims = oracle.dot_one_vs_db(g["cf_out"][2], g["cf_out"])
    idx = oracle.topk_desc(sims, 5)
    assert np.array_equal(idx, g["rt_top_idx"])
    assert np.array_equal(g["rt_labels"][idx], g["rt_top_labels"])
    np.testing.assert_allclose(sims[idx].astype(np.float64), g["rt_top_scores"], atol=TOL)


def test_votes_and_metrics(golden, oracle):
    g = golden("primitives.npz")
    vl, vs, true = g["v_labels"], g["v_scores"], g["v_true"]
    maj = np.array([oracle.vote_majority(list(r)) for r in vl])
    wv = np.array([oracle.vote_weighted(list(r), list(s), "f64") for r, s in zip(vl, vs)])
    assert float(np.mean(maj == true)) == float(g["v_acc_major"])
    assert float(np.mean(wv == true)) == float(g["v_acc_weight"])
    prf = oracle.prf_per_class(g["m_pred"], g["m_truth"], 4)
    got = np.array([[p["precision"], p["recall"], p["f1"], p["support"]] for p in prf])
    np.testing.assert_allclose(got, g["m_prf"], atol=1e-12)
    assert np.array_equal(oracle.confusion_counts(g["m_pred"], g["m_truth"], 4), g["m_cm"])


def test_topk_desc_edges(oracle):
    s = np.array([0.5, 0.9, 0.9, -1.0, 0.5], dtype=np.float32)
    assert list(oracle.topk_desc(s, 3)) == [1, 2, 0]          # ties -> lower index first
    assert list(oracle.topk_desc(s, 9)) == [1, 2, 0, 4, 3]    # k > n -> all
    assert list(oracle.topk_desc(s, 1)) == [1]


def test_cv_fold(golden, oracle):
    g = golden("cv_small.npz")
    n, d_img, d_txt, n_cls, pca_dim, top_k = [int(x) for x in g["meta"]]
    for f in range(5):
        tr, te = g[f"f{f}_train_idx"], g[f"f{f}_test_idx"]
        for fusion, w in (("concat", 0.5), ("late", 0.3), ("image_only", 0.5), ("text_only", 0.5)):
            r = oracle.cv_fold_eval(g[f"f{f}_img_tr"], g[f"f{f}_txt_tr"], g[f"f{f}_img_te"], g[f"f{f}_txt_te"],
                                    g["labels"][tr], g["labels"][te], n_cls, fusion=fusion, top_k=top_k,
                                    top_k_list=(1, 3, 5, 5), w_text=w)
            key = f"f{f}_{fusion}"
            assert np.array_equal(r["top_idx"], g[key + "_top_idx"]), key
            np.testing.assert_allclose(r["top_scores"], g[key + "_top_scores"], atol=TOL)
            got = np.array([r["top1"], r["top3"], r["top5"], r["vote_acc"], r["weighted_vote_acc"],
                            r["macro_precision"], r["macro_recall"], r["macro_f1"]])
            np.testing.assert_allclose(got, g[key + "_metrics"], atol=1e-12)
            assert np.array_equal(r["confusion_top1"], g[key + "_cm_top1"])
            assert np.array_equal(r["confusion_vote"], g[key + "_cm_vote"])
            assert np.array_equal(g["labels"][tr][r["top_idx"]], g[key + "_top_labels"])


def test_holdout(golden, oracle):
    g = golden("holdout_small.npz")
    args = (g["tr_txt"], g["te_txt"], g["tr_img"], g["te_img"], g["tr_labels"], g["te_labels"])
    runs = {
        "early": dict(fusion_type="early", text_weight=0.4),
        "late_none": dict(fusion_type="late", text_weight=0.4, score_mode="none"),
        "late_zscore": dict(fusion_type="late", text_weight=0.3, score_mode="zscore"),
        "late_minmax": dict(fusion_type="late", text_weight=0.6, score_mode="minmax"),
    }
    for name, kw in runs.items():
        r = oracle.holdout_eval(*args, top_k_list=(1, 3, 5, 7), **kw)
        for k, v in zip(g[name + "_keys"], g[name + "_vals"]):
            assert r[str(k)] == v, (name, k)
        if name + "_top5" in g:
            assert np.array_equal(np.array(r["all_top_labels_top5"]), g[name + "_top5"])
    r = oracle.holdout_eval(None, None, g["tr_img"], g["te_img"], g["tr_labels"], g["te_labels"],
                            fusion_type="none", top_k_list=(1, 3, 5, 5))
    for k, v in zip(g["imgonly_keys"], g["imgonly_vals"]):
        assert r[str(k)] == v
    sc = g["fs_scores"]
    assert oracle._scores_topk_acc(sc, g["tr_labels"], g["te_labels"], 3) == float(g["fs_top3"])
    assert oracle._scores_weighted_acc(sc, g["tr_labels"], g["te_labels"]) == float(g["fs_weighted"])


def test_batched_search_matches_loop(oracle):
    rng = np.random.default_rng(3)
    db = oracle.unit_rows(rng.standard_normal((500, 32)).astype(np.float32))
    db[17] = db[400]                                           # exact tie inside the top-k of query 400
    qs = db[395:405].copy()
    idx, sc = oracle.search_topk_batched(qs, db, 6)
    for i in range(len(qs)):
        ref = oracle.topk_desc(oracle.dot_one_vs_db(qs[i], db), 6)
        # sgemm vs sgemv may reorder exact ties only when rounding differs; compare as sets + scores
        assert set(idx[i]) == set(ref)
    fold_db = rng.integers(0, 5, size=500).astype(np.uint8)
    idx, sc = oracle.search_topk_batched(qs, db, 6, q_fold=fold_db[395:405], db_fold=fold_db)
    for i in range(len(qs)):
        assert np.all(fold_db[idx[i]] != fold_db[395 + i])


# ------------------------------------------------------------------ per-fold preprocessing
def test_scaler_oracle_equals_sklearn(oracle):
    from sklearn.preprocessing import StandardScaler
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((1000, 64)) * rng.uniform(0.01, 30, 64) + rng.uniform(-50, 50, 64)).astype(np.float32)
    x[:, 5] = 3.25                                   # constant feature -> scale 1
    sk = StandardScaler().fit(x)
    mean, scale = oracle.scaler_fit(x)
    np.testing.assert_allclose(mean, sk.mean_, rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(scale, sk.scale_, rtol=1e-13)
    assert scale[5] == 1.0
    assert np.array_equal(oracle.scaler_apply(x, mean, scale), sk.transform(x))


def test_exact_pca_oracle_is_what_sklearn_computes_in_float64(golden, oracle):
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import StandardScaler
    g = golden("cv_small.npz")
    z = StandardScaler().fit_transform(g["image"][g["f3_train_idx"]])
    comps, mean = oracle.pca_exact_fit(z, 16)
    for solver in ("full", "covariance_eigh"):
        sk = PCA(16, svd_solver=solver).fit(z.astype(np.float64))
        assert np.max(np.abs(sk.components_ - comps)) < 1e-10
        assert np.max(np.abs(sk.mean_ - mean)) < 1e-12


def test_process_embeddings_oracle_vs_reference_outputs(golden, oracle):
    """The reference's process_embeddings on the golden folds ran sklearn's fp32 "full" SVD (240 x 48 / 240 x 40):
    it sits within 2.5e-4 of the exact basis (fp32 LAPACK rounding at eigenvalue gaps of ~0.2 %)."""
    g = golden("cv_small.npz")
    pca_dim = int(g["meta"][4])
    worst = 0.0
    for f in range(5):
        tr_i, te_i = g[f"f{f}_train_idx"], g[f"f{f}_test_idx"]
        for mod, key in (("image", "img"), ("text", "txt")):
            tr, te = oracle.process_embeddings_exact(g[mod][tr_i], g[mod][te_i], pca_dim)
            assert tr.dtype == np.float32 and tr.shape == g[f"f{f}_{key}_tr"].shape
            worst = max(worst, np.abs(tr - g[f"f{f}_{key}_tr"]).max(), np.abs(te - g[f"f{f}_{key}_te"]).max())
    assert worst < 5e-4


def test_sklearn_solver_rule_matches_sklearn():
    from sklearn.decomposition import PCA
    from emr2a_b200.preprocess import sklearn_solver
    rng = np.random.default_rng(1)
    for n, d, p in ((240, 48, 16), (600, 40, 16), (520, 300, 16), (520, 300, 280), (100, 600, 50), (5200, 512, 128)):
        x = rng.standard_normal((n, d)).astype(np.float32)
        sk = PCA(n_components=min(p, n - 1, d)).fit(x)
        assert sklearn_solver(n, d, min(p, n - 1, d)) == sk._fit_svd_solver, (n, d, p)


@pytest.mark.parametrize("mode", ["zscore", "minmax"])
def test_late_fusion_affine_identity_behind_the_matrix_free_path(oracle, mode):
    """The derivation emr2a_b200/late.py relies on, checked in numpy: per query the reference's
    w * norm(ts) + (1 - w) * norm(is) equals <[g_t Tq ; g_i Iq], [Td ; Id]> - c with the statistics taken from
    database moments (z-score: column sums + Gram matrix) or from the extreme scores (min-max)."""
    rng = np.random.default_rng(3)
    n, d_t, d_i, w = 4000, 24, 40, 0.35
    db_t = oracle.unit_rows(rng.standard_normal((n, d_t)).astype(np.float32) + 0.4)
    db_i = oracle.unit_rows(rng.standard_normal((n, d_i)).astype(np.float32))
    s_t, g_t = db_t.astype(np.float64).sum(axis=0), db_t.astype(np.float64).T @ db_t.astype(np.float64)
    s_i, g_i = db_i.astype(np.float64).sum(axis=0), db_i.astype(np.float64).T @ db_i.astype(np.float64)
    for _ in range(5):
        q_t = oracle.unit_rows(rng.standard_normal((1, d_t)).astype(np.float32))[0]
        q_i = oracle.unit_rows(rng.standard_normal((1, d_i)).astype(np.float32))[0]
        ts, is_ = oracle.dot_one_vs_db(q_t, db_t), oracle.dot_one_vs_db(q_i, db_i)
        want = oracle.fuse_late_scores(ts, is_, w, mode)
        stats = []
        for q, s, g, sc in ((q_t, s_t, g_t, ts), (q_i, s_i, g_i, is_)):
            q64 = q.astype(np.float64)
            if mode == "zscore":
                mean = q64 @ s / n
                std = np.sqrt(max(q64 @ g @ q64 / n - mean * mean, 0.0))
                assert abs(mean - float(sc.mean())) < 1e-7 and abs(std - float(sc.std())) < 1e-7
                stats.append((mean, std + 1e-8))
            else:
                stats.append((float(sc.min()), float(sc.max()) - float(sc.min()) + 1e-8))
        (a_t, b_t), (a_i, b_i) = stats
        gam_t, gam_i = w / b_t, (1 - w) / b_i
        fused = np.concatenate([gam_t * q_t, gam_i * q_i]).astype(np.float64) @ np.concatenate([db_t, db_i], axis=1).astype(np.float64).T
        fused -= gam_t * a_t + gam_i * a_i
        assert np.max(np.abs(fused - want)) < 2e-5 * max(1.0, float(np.abs(want).max()))
        assert np.array_equal(np.argsort(-fused)[:5], np.argsort(-want.astype(np.float64))[:5])


@pytest.mark.parametrize("fusion", ["concat", "late", "image_only"])
def test_streamed_reference_sample_matches_reference_folds(golden, oracle, fusion):
    """The chunk-streamed sample loop (used for parity at the 1M-10M row sizes of C2-C5) reproduces the reference's
    own evaluate_fold outputs: Top-K rows, scores, labels, votes -- fed in ragged chunks."""
    g = golden("cv_small.npz")
    n_folds = 5
    k = int(g["meta"][5])
    for f in range(n_folds):
        tr, te = g[f"f{f}_train_idx"], g[f"f{f}_test_idx"]
        a_tr, a_te, b_tr, b_te = (g[f"f{f}_{nm}"] for nm in ("img_tr", "img_te", "txt_tr", "txt_te"))
        mode = {"concat": "concat", "late": "late", "image_only": "single"}[fusion]
        s = oracle.StreamedReferenceSample(mode, k, len(tr), a_te, None if mode == "single" else b_te,
                                           w_text=0.3 if fusion == "late" else 0.5)      # make_golden.py:150
        for r0 in range(0, len(tr), 37):
            s.add_chunk(r0, a_tr[r0:r0 + 37], None if mode == "single" else b_tr[r0:r0 + 37])
        r = s.finish(g["labels"][tr], g["labels"][te])
        want_idx, want_sc = g[f"f{f}_{fusion}_top_idx"], g[f"f{f}_{fusion}_top_scores"]
        par = oracle.sample_parity({"top_idx": want_idx, "top_scores": want_sc}, r["top_idx"], r["top_scores"])
        assert par["ok"] and par["max_score_err"] < TOL, par
        clear = np.abs(np.diff(want_sc, axis=1)).min(axis=1) > 1e-6
        assert np.array_equal(r["top_idx"][clear], want_idx[clear])
        assert np.array_equal(r["top_labels"][clear], g[f"f{f}_{fusion}_top_labels"][clear])
        metrics = g[f"f{f}_{fusion}_metrics"]         # top1, top3, top5, vote_acc, weighted_vote_acc, ...
        if clear.all():
            assert abs(r["top1"] - metrics[0]) < 1e-12 and abs(r["vote_acc"] - metrics[3]) < 1e-12
            assert abs(r["weighted_vote_acc"] - metrics[4]) < 1e-12


def test_streamed_reference_sample_fold_rule(oracle):
    """Own-fold rows are never scored (utils/cv_evaluator.py:349-376) and the result equals the brute-force masked
    ranking; the parity helper flags a wrong row on a clear gap and tolerates a swap inside the tolerance."""
    rng = np.random.default_rng(5)
    n, d, k = 3000, 48, 5
    img = rng.standard_normal((n, d), dtype=np.float32)
    txt = rng.standard_normal((n, d), dtype=np.float32)
    lab = rng.integers(0, 3, n)
    fold = np.sort(rng.integers(0, 5, n)).astype(np.uint8)
    qi = np.arange(7, n, 173)
    s = oracle.StreamedReferenceSample("concat", k, n, img[qi], txt[qi], q_fold=fold[qi])
    for r0 in range(0, n, 500):
        s.add_chunk(r0, img[r0:r0 + 500], txt[r0:r0 + 500], fold[r0:r0 + 500])
    r = s.finish(lab, lab[qi])
    db = oracle.fuse_concat_cv(oracle.unit_rows(img), oracle.unit_rows(txt))
    want_idx, want_sc = oracle.search_topk_batched(db[qi], db, k, q_fold=fold[qi], db_fold=fold)
    assert np.array_equal(r["top_idx"], want_idx) and np.max(np.abs(r["top_scores"] - want_sc)) < TOL
    assert not np.any(fold[r["top_idx"]] == fold[qi][:, None])
    assert r["rows_scored"] == int(sum((fold != f).sum() for f in fold[qi]))
    assert oracle.sample_parity(r, want_idx, want_sc, r["pred_vote"], r["pred_weighted"])["ok"]
    bad = want_idx.copy(); bad[3, 0] = (bad[3, 0] + 1) % n
    assert not oracle.sample_parity(r, bad, want_sc)["ok"]
    with pytest.raises(ValueError):
        oracle.StreamedReferenceSample("concat", k, n + 1, img[qi], txt[qi]).finish(lab)


def test_streamed_feed_timed_and_untimed_chunks_agree(oracle):
    """oracle/streamed_feed.feed: a bounded timed prefix (per-query sgemv) + thread-pooled sgemm chunks give the same
    ranking as feeding everything the reference's way, and the timings are scaled to the whole database."""
    import streamed_feed
    rng = np.random.default_rng(9)
    n, d, k = 4000, 32, 5
    img = rng.standard_normal((n, d), dtype=np.float32)
    txt = rng.standard_normal((n, d), dtype=np.float32)
    fold = np.sort(rng.integers(0, 5, n)).astype(np.uint8)
    lab = rng.integers(0, 3, n)
    qi = np.arange(3, n, 211)

    def fetch(r0, rows):
        return img[r0:r0 + rows], txt[r0:r0 + rows], fold[r0:r0 + rows]
    for mode in ("concat", "late"):
        a = oracle.StreamedReferenceSample(mode, k, n, img[qi], txt[qi], w_text=0.25, q_fold=fold[qi])
        streamed_feed.feed(a, n, fetch, chunk_rows=512, timed_rows=n)
        b = oracle.StreamedReferenceSample(mode, k, n, img[qi], txt[qi], w_text=0.25, q_fold=fold[qi])
        streamed_feed.feed(b, n, fetch, chunk_rows=512, timed_rows=1024, workers=3)
        ra, rb = a.finish(lab, lab[qi]), b.finish(lab, lab[qi])
        assert oracle.sample_parity(ra, rb["top_idx"], rb["top_scores"], rb["pred_vote"], rb["pred_weighted"], tol=1e-6)["ok"]
        assert ra["rows_scored"] == rb["rows_scored"] and rb["timed_rows"] == 1024 and ra["timed_rows"] == n
        assert rb["prep_seconds"] > rb["timed_prep_seconds"] > 0
