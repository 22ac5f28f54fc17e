"""GPU parity of the per-fold preprocessing: emr2a_column_moments / emr2a_standardize through
the C-ABI, the exact PCA of emr2a_b200.preprocess against the numpy oracle and against sklearn, and the
evaluators with ``preprocess = "gpu"`` against the reference outputs in tests/golden/cv_small.npz."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _structured(n, d, seed, decay=0.93, offset=5.0):
    """Rows with a geometrically decaying spectrum (well separated principal axes), shifted and rescaled columns."""
    rng = np.random.default_rng(seed)
    basis, _ = np.linalg.qr(rng.standard_normal((d, d)))
    lat = rng.standard_normal((n, d)) * (decay ** np.arange(d))[None, :]
    x = lat @ basis.T
    x = x * rng.uniform(0.2, 5.0, size=d)[None, :] + rng.uniform(-offset, offset, size=d)[None, :]
    return x.astype(np.float32)


@pytest.mark.parametrize("n,d,pad", [(1, 7, 0), (63, 40, 0), (1000, 130, 3), (4097, 512, 0), (20000, 96, 0)])
def test_column_moments_match_float64_numpy(n, d, pad):
    import torch
    from emr2a_b200 import preprocess as pp
    from emr2a_b200.engine import get_engine
    eng = get_engine()
    rng = np.random.default_rng(n + d)
    full = (rng.standard_normal((n, d + pad)) * 3 + 11).astype(np.float32)
    x = torch.from_numpy(full).to(eng.device)[:, :d]            # leading dimension d + pad (unaligned when pad = 3)
    shift = torch.from_numpy(full[0, :d].copy()).to(eng.device)
    for sh in (None, shift):
        s, ss = pp.column_moments(eng, x, sh)
        ref = full[:, :d].astype(np.float64) - (0.0 if sh is None else full[0, :d].astype(np.float64))
        np.testing.assert_allclose(s.cpu().numpy(), ref.sum(axis=0), rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(ss.cpu().numpy(), (ref ** 2).sum(axis=0), rtol=1e-12)
    s2, ss2 = pp.column_moments(eng, x, shift)                   # deterministic: bit-identical on repeat
    assert torch.equal(s, s2) and torch.equal(ss, ss2)


@pytest.mark.parametrize("n,d", [(5, 3), (777, 48), (5000, 512), (300, 1030)])
def test_scaler_is_bit_identical_to_sklearn(n, d):
    from sklearn.preprocessing import StandardScaler
    from emr2a_b200 import preprocess as pp
    from emr2a_b200.engine import get_engine
    eng = get_engine()
    rng = np.random.default_rng(7 * n + d)
    x = (rng.standard_normal((n, d)) * rng.uniform(0.01, 30, d) + rng.uniform(-50, 50, d)).astype(np.float32)
    x[:, d // 2] = -2.5                                            # constant feature
    y = (rng.standard_normal((n // 2 + 1, d)) * 4).astype(np.float32)
    sk = StandardScaler().fit(x)
    tf = pp.fit(x, None, eng)
    np.testing.assert_allclose(tf.mean.cpu().numpy(), sk.mean_, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(tf.scale.cpu().numpy(), sk.scale_, rtol=1e-12)
    assert float(tf.scale[d // 2]) == 1.0
    for arr in (x, y):
        got = pp.transform(tf, arr, eng, normalize=False).cpu().numpy()
        want = sk.transform(arr)
        assert np.mean(got != want) < 1e-3                         # mean_/scale_ can differ in the last float64 bit
        np.testing.assert_allclose(got, want, rtol=3e-7, atol=1e-7)


@pytest.mark.parametrize("n,d,p", [(240, 48, 16), (1600, 512, 128), (60, 200, 128), (6000, 64, 64), (300, 40, 1)])
def test_exact_pca_matches_oracle(oracle, n, d, p):
    from emr2a_b200 import preprocess as pp
    from emr2a_b200.engine import get_engine
    eng = get_engine()
    x = _structured(n + 100, d, seed=n + p)
    tr, te = x[:n], x[n:]
    tf = pp.fit(tr, p, eng)
    assert tf.n_components == min(p, n - 1, d)
    o_tr, o_te = oracle.process_embeddings_exact(tr, te, p)
    g_tr = pp.transform(tf, tr, eng).cpu().numpy()
    g_te = pp.transform(tf, te, eng).cpu().numpy()
    assert g_tr.shape == o_tr.shape and g_tr.dtype == np.float32
    # rank-deficient folds (n <= d): the trailing axes carry only rounding noise; compare the well-determined ones
    keep = min(tf.n_components, n - 2)
    assert np.max(np.abs(g_tr[:, :keep] - o_tr[:, :keep])) < 2e-5
    assert np.max(np.abs(g_te[:, :keep] - o_te[:, :keep])) < 2e-5


def test_exact_pca_against_sklearn_solvers():
    """sklearn's deterministic solvers approximate the same basis in fp32."""
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import StandardScaler
    from emr2a_b200 import preprocess as pp
    from emr2a_b200.engine import get_engine
    eng = get_engine()
    x = _structured(6000, 64, seed=5, decay=0.9)
    tf = pp.fit(x, 24, eng)
    z = StandardScaler().fit_transform(x)
    for solver, tol in (("covariance_eigh", 2e-4), ("full", 2e-3)):
        sk = PCA(24, svd_solver=solver).fit(z)
        assert np.max(np.abs(sk.components_ - tf.components.cpu().numpy())) < tol, solver
        got = pp.transform(tf, x[:500], eng, normalize=False).cpu().numpy()
        assert np.max(np.abs(sk.transform(z[:500]) - got)) < 20 * tol, solver
    np.testing.assert_allclose(tf.explained_variance.cpu().numpy(),
                               PCA(24, svd_solver="full").fit(z.astype(np.float64)).explained_variance_, rtol=1e-9)


def test_chunked_fit_and_transform_equal_single_chunk(monkeypatch):
    import torch
    from emr2a_b200 import preprocess as pp
    from emr2a_b200.engine import get_engine
    eng = get_engine()
    x = _structured(5000, 72, seed=9)
    tf_a = pp.fit(x, 20, eng)
    y_a = pp.transform(tf_a, x, eng)
    monkeypatch.setattr(pp, "_CHUNK_ROWS", 1024)
    tf_b = pp.fit(x, 20, eng)
    y_b = pp.transform(tf_b, x, eng)
    assert torch.equal(tf_a.mean, tf_b.mean) and torch.equal(tf_a.scale, tf_b.scale)
    assert float((tf_a.components - tf_b.components).abs().max()) < 1e-6
    assert float((y_a - y_b).abs().max()) < 5e-6


def test_cv_evaluator_gpu_preprocessing_on_golden_folds(golden, oracle):
    from emr2a_b200.utils.cv_evaluator import CVRetrievalEvaluator
    g = golden("cv_small.npz")
    n, d_img, d_txt, n_cls, pca_dim, top_k = [int(v) for v in g["meta"]]
    ev = CVRetrievalEvaluator(cv_folds=5, pca_dim=pca_dim, top_k=top_k, seed=42)
    ev.preprocess = "gpu"
    names = [f"class_{c}" for c in g["labels"]]
    for f in range(5):
        tr_i, te_i = g[f"f{f}_train_idx"], g[f"f{f}_test_idx"]
        a_tr, a_te = ev.process_embeddings(g["image"][tr_i], g["image"][te_i])
        assert isinstance(a_tr, np.ndarray) and a_tr.dtype == np.float32
        # reference output (sklearn fp32 "full" SVD) within its own rounding of the exact basis
        assert np.max(np.abs(a_tr - g[f"f{f}_img_tr"])) < 5e-4 and np.max(np.abs(a_te - g[f"f{f}_img_te"])) < 5e-4
        o_tr, o_te = oracle.process_embeddings_exact(g["image"][tr_i], g["image"][te_i], pca_dim)
        assert np.max(np.abs(a_tr - o_tr)) < 1e-5 and np.max(np.abs(a_te - o_te)) < 1e-5
        # whole fold: exact-basis oracle pipeline, Top-K rows identical where the score gaps are clear
        b_tr, b_te = oracle.process_embeddings_exact(g["text"][tr_i], g["text"][te_i], pca_dim)
        tr_lab, te_lab = [names[j] for j in tr_i], [names[j] for j in te_i]
        r = ev.evaluate_fold(g["image"][tr_i], g["text"][tr_i], g["image"][te_i], g["text"][te_i], tr_lab, te_lab,
                             [f"p{j}" for j in te_i], fusion="concat", top_k_list=[1, 3, 5, 5],
                             train_ids=[f"p{j}" for j in tr_i])
        db = oracle.fuse_concat_cv(o_tr, b_tr)
        qs = oracle.fuse_concat_cv(o_te, b_te)
        o_idx, o_sc = oracle.search_topk_batched(qs, db, top_k)
        got_sc = np.array(r["all_top_scores"])
        assert np.max(np.abs(got_sc - o_sc)) < 2e-5
        clear = np.abs(np.diff(o_sc, axis=1)).min(axis=1) > 1e-4
        got_idx = np.array([[int(p[1:]) for p in row] for row in r["all_top_patient_ids"]])
        assert clear.sum() > 30 and np.array_equal(got_idx[clear], tr_i[o_idx[clear]])
        # and the reference's metrics of the same fold (its PCA differs by 2e-4: a few near-tie flips at most)
        ref = g[f"f{f}_concat_metrics"]
        got = np.array([r["top1"], r["top3"], r["top5"], r["vote_acc"], r["weighted_vote_acc"]])
        assert np.max(np.abs(got - ref[:5])) <= 2.0 / len(te_i) + 1e-12


def test_holdout_evaluator_gpu_scaler_equals_host(oracle):
    from emr2a_b200.retrieval.evaluator import RetrievalEvaluator
    x = _structured(400, 48, seed=3)
    ho = RetrievalEvaluator(use_pca=False)
    ho.preprocess = "host"
    h_tr, h_te = ho.process_embeddings(x[:300], x[300:])
    ho.preprocess = "gpu"
    g_tr, g_te = ho.process_embeddings(x[:300], x[300:])
    assert np.max(np.abs(h_tr - g_tr)) < 3e-7 and np.max(np.abs(h_te - g_te)) < 3e-7
    ho = RetrievalEvaluator(use_pca=True, pca_dim=12)
    ho.preprocess = "gpu"
    g_tr, g_te = ho.process_embeddings(x[:300], x[300:])
    o_tr, o_te = oracle.process_embeddings_exact(x[:300], x[300:], 12)
    assert np.max(np.abs(g_tr - o_tr)) < 1e-5 and np.max(np.abs(g_te - o_te)) < 1e-5


@pytest.mark.parametrize("fusion,w", [("concat", 0.5), ("late", 0.3), ("image_only", 0.5)])
def test_run_cv_arrays_equals_run_cv_with_gpu_preprocessing(golden, fusion, w):
    """The array form of the whole pipeline against the dict-shaped run_cv (both with device preprocessing)."""
    from emr2a_b200 import synth
    from emr2a_b200.utils.cv_evaluator import CVRetrievalEvaluator
    g = golden("cv_small.npz")
    n, d_img, d_txt, n_cls, pca_dim, top_k = [int(v) for v in g["meta"]]
    ids = synth.patient_ids(n)
    labels = [f"class_{c}" for c in g["labels"]]
    emb = {pid: {"image": g["image"][j], "text": g["text"][j]} for j, pid in enumerate(ids)}
    ev = CVRetrievalEvaluator(cv_folds=5, pca_dim=pca_dim, top_k=top_k, seed=42)
    ev.preprocess = "gpu"
    a = ev.run_cv(ids, labels, emb, fusion=fusion, top_k_list=[1, 3, 5, 5], w_text=w)
    b = ev.run_cv_arrays(labels, g["image"], g["text"], fusion=fusion, top_k_list=[1, 3, 5, 5], w_text=w,
                         patient_ids=ids, lists=True)
    c = ev.run_cv_arrays(np.asarray(g["labels"]), g["image"], g["text"], fusion=fusion, top_k_list=[1, 3, 5, 5], w_text=w)
    assert a["summary"] == b["summary"]
    for ra, rb, rc in zip(a["fold_results"], b["fold_results"], c["fold_results"]):
        assert ra["test_patient_ids"] == rb["test_patient_ids"] and ra["train_ids"] == rb["train_ids"]
        assert ra["all_top_patient_ids"] == rb["all_top_patient_ids"] and ra["all_top_labels"] == rb["all_top_labels"]
        assert ra["all_top_scores"] == rb["all_top_scores"]
        assert ra["confusion_matrix_vote"] == rb["confusion_matrix_vote"]
        assert "all_top_labels" not in rc and rc["train_ids"] == []
        for key in ("top1", "top3", "top5", "vote_acc", "weighted_vote_acc", "macro_f1"):
            assert ra[key] == rc[key]
    with pytest.raises(ValueError):
        ev.run_cv_arrays(labels, g["image"], None, fusion="concat")
    with pytest.raises(ValueError):
        ev.run_cv_arrays(labels, g["image"], g["text"], fusion="bogus")


@pytest.mark.parametrize("n,d,pad", [(1, 5, 0), (130, 64, 0), (1000, 130, 3), (4097, 512, 0), (50000, 96, 0), (300, 1030, 2)])
def test_gram_f64_with_fused_standardisation(n, d, pad):
    """emr2a_gram_f64: Z^T Z and column sums of Z in float64, Z standardised on the fly in sklearn's fp32 arithmetic --
    against float64 numpy on the materialised fp32 Z (emr2a_standardize), with and without the scaler; symmetric,
    deterministic."""
    import torch
    from emr2a_b200 import native, preprocess as pp
    from emr2a_b200.engine import get_engine
    eng = get_engine()
    rng = np.random.default_rng(3 * n + d)
    full = (rng.standard_normal((n, d + pad)) * rng.uniform(0.1, 9, d + pad) + rng.uniform(-20, 20, d + pad)).astype(np.float32)
    x = torch.from_numpy(full).to(eng.device)[:, :d]
    mean = torch.from_numpy(full[:, :d].mean(axis=0).astype(np.float32)).to(eng.device)
    scale = torch.from_numpy((full[:, :d].std(axis=0) + 0.1).astype(np.float32)).to(eng.device)

    def gram(m, s):
        g = torch.empty((d, d), dtype=torch.float64, device=eng.device)
        zs = torch.empty((d,), dtype=torch.float64, device=eng.device)
        ws = torch.empty((int(eng.lib.emr2a_gram_f64_workspace_bytes(n, d)) // 8 + 2,), dtype=torch.float64, device=eng.device)
        native.check(eng.lib.emr2a_gram_f64(x.data_ptr(), int(x.stride(0)) if n > 1 else d, n, d, native.ptr(m), native.ptr(s),
                                            g.data_ptr(), zs.data_ptr(), ws.data_ptr(), ws.numel() * 8, None))
        return g, zs
    for m, s in ((None, None), (mean, scale)):
        z = full[:, :d] if m is None else pp.standardize(eng, x, m, s).cpu().numpy()
        z64 = z.astype(np.float64)
        g, zs = gram(m, s)
        want = z64.T @ z64
        np.testing.assert_allclose(g.cpu().numpy(), want, rtol=1e-12, atol=1e-9 * np.abs(want).max())
        np.testing.assert_allclose(zs.cpu().numpy(), z64.sum(axis=0), rtol=1e-12, atol=1e-9 * max(1.0, np.abs(z64).sum(axis=0).max()))
        assert torch.equal(g, g.t())
        g2, zs2 = gram(m, s)
        assert torch.equal(g, g2) and torch.equal(zs, zs2)


@pytest.mark.parametrize("n,d,p", [(1, 8, 3), (777, 48, 16), (5000, 512, 128), (300, 1030, 200), (70000, 64, 24)])
def test_project_equals_standardize_scores_bias(n, d, p):
    """emr2a_project (standardisation fused into the operand load, bias into the epilogue) is bit-identical to the
    three separate steps emr2a_standardize -> emr2a_scores -> subtract."""
    import torch
    from emr2a_b200 import native, preprocess as pp
    from emr2a_b200.engine import get_engine
    eng = get_engine()
    g = torch.Generator(device=eng.device).manual_seed(n + d + p)
    x = torch.randn((n, d), generator=g, device=eng.device) * 3 + 1
    mean = torch.randn((d,), generator=g, device=eng.device)
    scale = torch.rand((d,), generator=g, device=eng.device) + 0.5
    w = torch.randn((p, d), generator=g, device=eng.device) / d ** 0.5
    bias = torch.randn((p,), generator=g, device=eng.device)
    want = eng.scores(pp.standardize(eng, x, mean, scale), w) - bias
    got = torch.empty((n, p), dtype=torch.float32, device=eng.device)
    native.check(eng.lib.emr2a_project(x.data_ptr(), d, n, d, mean.data_ptr(), scale.data_ptr(), w.data_ptr(), d, p,
                                       bias.data_ptr(), got.data_ptr(), p, None))
    assert torch.equal(got, want)
    native.check(eng.lib.emr2a_project(x.data_ptr(), d, n, d, None, None, w.data_ptr(), d, p, None, got.data_ptr(), p, None))
    assert torch.equal(got, eng.scores(x, w))


@pytest.mark.parametrize("n,d", [(5, 8), (777, 48), (5000, 512), (3000, 1024), (200, 2048)])
def test_k1_fused_standardisation_equals_two_passes(n, d):
    """K1 with NF_STANDARDIZE (scaler + row normalisation in one pass over the raw rows) equals emr2a_standardize
    followed by K1: the per-column division is the correctly rounded quotient in both."""
    import torch
    from emr2a_b200 import native, preprocess as pp
    from emr2a_b200.engine import get_engine
    eng = get_engine()
    g = torch.Generator(device=eng.device).manual_seed(n * 7 + d)
    x = torch.randn((n, d), generator=g, device=eng.device) * 4 - 2
    mean = torch.randn((d,), generator=g, device=eng.device)
    scale = torch.rand((d,), generator=g, device=eng.device) * 3 + 0.05
    z = pp.standardize(eng, x, mean, scale)
    want = eng.normalize_fuse(z, flags=native.NF_ROWNORM).f32
    col_std = torch.stack([mean, scale, 1.0 / scale]).contiguous()
    got = eng.normalize_fuse(x, flags=native.NF_ROWNORM | native.NF_STANDARDIZE, col_std=col_std).f32
    assert torch.equal(got, want)
    # without the row normalisation the kernel is StandardScaler.transform itself
    assert torch.equal(eng.normalize_fuse(x, flags=native.NF_STANDARDIZE, col_std=col_std).f32, z)
    # shapes the fused variant does not take are refused loudly (the host layer then runs the two passes)
    with pytest.raises(native.Emr2aError):
        eng.normalize_fuse(x[:, :d - 1].contiguous(), flags=native.NF_ROWNORM | native.NF_STANDARDIZE,
                           col_std=col_std[:, :d - 1].contiguous())
    tf = pp.fit(x, None, eng)
    a = pp.transform(tf, x, eng)                                   # scaler-only transform takes the fused pass
    b = eng.normalize_fuse(pp.standardize(eng, x, tf.mean_f32, tf.scale_f32), flags=native.NF_ROWNORM).f32
    assert torch.equal(a, b)
