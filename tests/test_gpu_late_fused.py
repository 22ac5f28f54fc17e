"""Late fusion with z-score / min-max normalisation without the score matrix: the fused search of
emr2a_b200/late.py against the oracle's materialised, op-for-op restatement of retrieval/fusion.py:4-14,31-42."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MODES = {"none": 0, "zscore": 1, "minmax": 2}


def _data(n_db, n_q, d_t, d_i, seed):
    from emr2a_b200 import synth
    both = synth.two_modal(n_db + n_q, d_i, d_t, 3, seed=seed, sep=0.3)
    scale = np.random.default_rng(seed).uniform(0.3, 3.0, size=(n_db + n_q, 1)).astype(np.float32)
    txt, img = both["text"] * scale, both["image"] / scale
    return (txt[:n_db], img[:n_db], txt[n_db:], img[n_db:], both["labels"][:n_db], both["labels"][n_db:])


def _oracle_topk(oracle, db_t, db_i, q_t, q_i, w, mode, k):
    idx, sc = [], []
    for j in range(len(q_t)):
        st = oracle.cosine_one_vs_db(q_t[j], db_t)
        si = oracle.cosine_one_vs_db(q_i[j], db_i)
        fused = oracle.fuse_late_scores(st, si, w, mode)
        top = oracle.topk_desc(fused, k)
        idx.append(top)
        sc.append(fused[top])
    return np.array(idx), np.array(sc)


@pytest.mark.parametrize("mode,d_t,d_i", [("zscore", 72, 56), ("minmax", 72, 56), ("none", 72, 56),
                                          ("minmax", 64, 128)])       # 64 | 128: min-max on column-sliced operands
@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "rescore"])
def test_fused_late_search_matches_materialised_reference(oracle, mode, d_t, d_i, prec):
    from emr2a_b200.engine import get_engine, unpack_keys
    from emr2a_b200.late import late_fusion_search
    eng = get_engine()
    db_t, db_i, q_t, q_i, _, _ = _data(6000, 150, d_t, d_i, seed=41)
    k, w = 5, 0.35
    keys = late_fusion_search(db_t, db_i, q_t, q_i, w, MODES[mode], k, precision=prec, engine=eng)
    sc, idx = unpack_keys(keys)
    o_idx, o_sc = _oracle_topk(oracle, db_t, db_i, q_t, q_i, w, mode, k)
    # z-scores are O(1..5): 1e-5 relative to that magnitude (fp32 chains of different length on both sides)
    tol = 1e-5 * max(1.0, float(np.abs(o_sc).max()))
    assert np.max(np.abs(sc - o_sc)) < tol, float(np.max(np.abs(sc - o_sc)))
    clear = np.abs(np.diff(o_sc, axis=1)).min(axis=1) > 4 * tol
    assert clear.sum() > 100
    assert np.array_equal(idx[clear], o_idx[clear])


def test_fused_late_search_degenerate_queries(oracle):
    """A zero text query (all text scores 0 -> std 0, range 0: the epsilon path) and a database of duplicates."""
    from emr2a_b200.engine import get_engine, unpack_keys
    from emr2a_b200.late import late_fusion_search
    eng = get_engine()
    db_t, db_i, q_t, q_i, _, _ = _data(500, 8, 16, 24, seed=2)
    q_t[0] = 0.0
    for mode in ("zscore", "minmax"):
        keys = late_fusion_search(db_t, db_i, q_t, q_i, 0.4, MODES[mode], 3, precision="fp32", engine=eng)
        sc, idx = unpack_keys(keys)
        o_idx, o_sc = _oracle_topk(oracle, db_t, db_i, q_t, q_i, 0.4, mode, 3)
        assert np.all(np.isfinite(sc))
        assert np.max(np.abs(sc - o_sc)) < 3e-5 * max(1.0, float(np.abs(o_sc).max()))
        assert np.array_equal(idx[0], o_idx[0])


@pytest.mark.parametrize("mode", ["zscore", "minmax"])
def test_holdout_evaluator_fused_equals_materialised(monkeypatch, oracle, mode):
    """RetrievalEvaluator.evaluate_retrieval(score_mode=...) through both paths and against the oracle."""
    from emr2a_b200.retrieval.evaluator import RetrievalEvaluator
    db_t, db_i, q_t, q_i, db_l, q_l = _data(3000, 200, 40, 32, seed=17)
    names = lambda codes: [f"class_{c}" for c in codes]                       # noqa: E731
    out = {}
    for forced in ("0", "1"):
        monkeypatch.setenv("EMR2A_LATE_FUSED", forced)
        out[forced] = RetrievalEvaluator().evaluate_retrieval(db_t, q_t, db_i, q_i, names(db_l), names(q_l),
                                                              text_weight=0.4, fusion_type="late", score_mode=mode,
                                                              top_k_list=[1, 3, 5])
    want = oracle.holdout_eval(db_t, q_t, db_i, q_i, db_l, q_l, text_weight=0.4, fusion_type="late", score_mode=mode,
                               top_k_list=[1, 3, 5])
    for key in ("top1", "top3", "top5", "weighted"):
        assert abs(out["1"][key] - out["0"][key]) <= 1.0 / 200 + 1e-12, key
        assert abs(out["1"][key] - want[key]) <= 1.0 / 200 + 1e-12, key
    same = np.mean([a == b for a, b in zip(out["1"]["all_top_labels_top5"], out["0"]["all_top_labels_top5"])])
    assert same > 0.97


def test_scale_segments_and_key_offset_kernels():
    import torch
    from emr2a_b200 import native
    from emr2a_b200.engine import get_engine, unpack_keys
    eng = get_engine()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((37, 29)).astype(np.float32)
    g0, g1 = rng.uniform(0.5, 2, 37).astype(np.float32), rng.uniform(0.5, 2, 37).astype(np.float32)
    xd = torch.from_numpy(x.copy()).to(eng.device)
    g0d, g1d = torch.from_numpy(g0).to(eng.device), torch.from_numpy(g1).to(eng.device)
    native.check(eng.lib.emr2a_scale_segments(xd.data_ptr(), 37, 12, 17, 29, g0d.data_ptr(), g1d.data_ptr(), eng._stream()))
    want = x.copy()
    want[:, :12] *= g0[:, None]
    want[:, 12:] *= g1[:, None]
    assert np.array_equal(xd.cpu().numpy(), want)
    q = eng.normalize_fuse(x[:5], flags=0)
    db = eng.normalize_fuse(x, flags=0)
    keys = eng.topk_search(q, db, 40, "fp32")                                 # K > N: trailing empty slots stay empty
    sc0, idx0 = unpack_keys(keys)
    off = torch.tensor([1.5, -2.0, 0.0, 100.0, -0.25], dtype=torch.float32, device=eng.device)
    native.check(eng.lib.emr2a_keys_add_offset(keys.data_ptr(), 5, 40, off.data_ptr(), eng._stream()))
    sc1, idx1 = unpack_keys(keys)
    assert np.array_equal(idx0, idx1) and np.all(idx1[:, 37:] == -1)
    assert np.array_equal(sc1[:, :37], sc0[:, :37] + off.cpu().numpy()[:, None])


def test_resident_late_index_reuses_moments_and_matches_one_shot(oracle):
    import torch
    from emr2a_b200.engine import get_engine
    from emr2a_b200.late import LateFusionIndex, late_fusion_search
    eng = get_engine()
    db_t, db_i, q_t, q_i, _, _ = _data(5000, 120, 40, 48, seed=8)
    index = LateFusionIndex(db_t, db_i, k=5, precision="fp32", engine=eng)
    first = index.search(q_t[:60], q_i[:60], 0.3, MODES["zscore"], 5)
    cached = index._moments
    second = index.search(q_t[60:], q_i[60:], 0.3, MODES["zscore"], 5)
    assert index._moments is cached
    whole = late_fusion_search(db_t, db_i, q_t, q_i, 0.3, MODES["zscore"], 5, precision="fp32", engine=eng)
    assert torch.equal(torch.cat([first, second]), whole)
    mm = index.search(q_t, q_i, 0.3, MODES["minmax"], 5)
    assert torch.equal(mm, late_fusion_search(db_t, db_i, q_t, q_i, 0.3, MODES["minmax"], 5, precision="fp32", engine=eng))
    with pytest.raises(ValueError):
        LateFusionIndex(db_t, db_i[:10], engine=eng)
