""""step3_retrieval and analysis/run_cv_experiments.py run unchanged": replay on the GPU.

tests/golden/make_script_goldens.py ran the reference's two entry points UNCHANGED (in the build container, where the
reference tree exists) on a small synthetic cohort, with recording wrappers around the evaluator classes the scripts
construct (pipelines/step3_retrieval/evaluate_retrieval.py:64-78; analysis/run_cv_experiments.py:383-397, 490-495), and
committed (a) every constructor / call argument it saw and (b) the files the reference wrote.  The reference tree does
not exist on the GPU box, so here exactly those calls are replayed on the drop-in classes and the files they write are
compared with the reference's, key for key and value for value, under the gap rule: lists identical on rows whose
adjacent reference scores are more than 2e-5 apart, scores within 1e-5, scalar metrics / confusion matrices /
summary.csv EXACT when every row of the fold is clear (else they may move by at most the near-tied queries)."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden", "scripts")
TOL = 1e-5


def _calls():
    with open(os.path.join(GOLD, "calls.json"), encoding="utf-8") as fh:
        calls = json.load(fh)
    arrays = dict(np.load(os.path.join(GOLD, "calls.npz"), allow_pickle=False))
    return calls, arrays


def _resolve(v, arrays):
    if isinstance(v, dict) and set(v) == {"npz"}:
        return arrays[v["npz"]]
    return v


def test_recorded_arguments_are_what_the_input_files_hold():
    """The recorded evaluator inputs are the scripts' own reading of the input files: slice means of the per-patient
    arrays (evaluate_retrieval.py:66-67 / run_cv_experiments.py:316-333), in the scripts' patient order -- and the
    drop-in's array ingest (emr2a_b200/ingest.py, GPU mean-pool) reproduces them bit for bit."""
    from emr2a_b200 import ingest
    calls, arrays = _calls()
    c = dict(np.load(os.path.join(GOLD, "inputs.npz"), allow_pickle=False))
    ids = [str(p) for p in c["ids"]]
    assert [str(p) for p in arrays["cv_concat_patient_ids"]] == ids
    want = c["image_slices"].mean(axis=1)
    assert np.array_equal(arrays["cv_concat_image"], want) and np.array_equal(arrays["cv_concat_text"], c["text"])
    pooled = ingest.mean_pool_patients([c["image_slices"][j] for j in range(len(ids))]).cpu().numpy()
    assert np.array_equal(pooled, want)
    assert calls["cv_concat"]["run_cv"]["labels"] == [f"class_{int(x)}" for x in c["labels"]]
    n_tr = len(calls["step3"]["evaluate_retrieval"]["train_labels"])
    assert arrays["step3_train_image"].shape == (n_tr, c["image_slices"].shape[2]) and n_tr == 120


def test_step3_retrieval_replay(tmp_path, oracle):
    """pipelines/step3_retrieval/evaluate_retrieval.py:64-86: RetrievalEvaluator() ->
    evaluate_retrieval(train_text=None, ..., fusion_type="none", top_k_list=[1, 3, 5, top_k]) -> json.dump(results)."""
    from emr2a_b200.retrieval import RetrievalEvaluator
    calls, arrays = _calls()
    rec = calls["step3"]
    ev = RetrievalEvaluator(*rec["ctor_args"], **rec["ctor_kwargs"])
    kw = {k: _resolve(v, arrays) for k, v in rec["evaluate_retrieval"].items()}
    results = ev.evaluate_retrieval(**kw)
    out = tmp_path / "retrieval_results.json"
    with out.open("w", encoding="utf-8") as fh:                    # evaluate_retrieval.py:83-86
        json.dump(results, fh, ensure_ascii=False, indent=2)
    got = json.load(open(out))
    want = json.load(open(os.path.join(GOLD, "step3", "retrieval_results.json")))
    assert list(got) == list(want)                                 # same keys in the same order
    # near-ties among the reference's cosine scores (oracle = retrieval/similarity.py:4-7 restated)
    tr, te = kw["train_image"], kw["test_image"]
    sc = np.array([oracle.cosine_one_vs_db(te[i], tr) for i in range(len(te))], dtype=np.float64)
    top = np.sort(sc, axis=1)[:, ::-1]
    for key in want:
        depth = 5 if key.endswith("weighted") else int(key.rsplit("top", 1)[1])
        unclear = int(((top[:, :depth] - top[:, 1:depth + 1]).min(axis=1) <= 2 * TOL).sum())
        assert abs(got[key] - want[key]) <= unclear / len(te) + 1e-12, (key, got[key], want[key], unclear)
    assert sum(got[k] == want[k] for k in want) >= 3               # (nearly) all of them exactly
    if out.read_text() != open(os.path.join(GOLD, "step3", "retrieval_results.json")).read():
        assert any(got[k] != want[k] for k in want)                # byte-identical unless a near-tie moved a metric


@pytest.mark.parametrize("tag", ["cv_concat", "cv_late"])
def test_run_cv_experiments_replay(tmp_path, tag):
    """analysis/run_cv_experiments.py:383-397 (CVRetrievalEvaluator(cv_folds=5, pca_dim, top_k, seed) -> run_cv(...))
    and :471-494 (save_results(results, output_dir, experiment_id, config)): same calls, same files."""
    from emr2a_b200.utils.cv_evaluator import CVRetrievalEvaluator
    calls, arrays = _calls()
    rec = calls[tag]
    ev = CVRetrievalEvaluator(*rec["ctor_args"], **rec["ctor_kwargs"])
    assert ev.preprocess in ("host", "auto")                       # sklearn preprocessing as the reference (seeded below)
    run = rec["run_cv"]
    ids = [str(p) for p in arrays[run["patient_ids"]["npz"]]]
    img, txt = arrays[run["embeddings"]["image"]["npz"]], arrays[run["embeddings"]["text"]["npz"]]
    emb = {pid: {"image": img[j], "text": txt[j]} for j, pid in enumerate(ids)}
    np.random.seed(run["numpy_seed_before_call"])
    results = ev.run_cv(patient_ids=ids, labels=run["labels"], embeddings=emb, fusion=run["fusion"],
                        top_k_list=run["top_k_list"], w_text=run["w_text"])
    ev.save_results(results=results, output_dir=tmp_path, experiment_id=rec["save_results"]["experiment_id"],
                    config=rec["save_results"]["config"])
    exp, gold = tmp_path / f"exp_{tag}", os.path.join(GOLD, tag)
    produced = sorted(p.relative_to(exp).as_posix() for p in exp.rglob("*") if p.is_file() and p.suffix != ".png")
    expected = sorted(os.path.relpath(os.path.join(r, f), gold) for r, _, fs in os.walk(gold) for f in fs)
    assert produced == expected
    assert json.load(open(exp / "config.json")) == json.load(open(os.path.join(gold, "config.json")))
    all_clear = True
    scalar = ("top1", "top3", "top5", "vote_acc", "weighted_vote_acc", "macro_precision", "macro_recall", "macro_f1")
    for f in range(1, 6):
        got = json.load(open(exp / f"fold_{f}" / "metrics.json"))
        want = json.load(open(os.path.join(gold, f"fold_{f}", "metrics.json")))
        assert sorted(got) == sorted(want), f
        for key in ("fold", "test_patient_ids", "train_ids"):
            assert got[key] == want[key], (f, key)
        w_sc, g_sc = np.array(want["all_top_scores"]), np.array(got["all_top_scores"])
        assert w_sc.shape == g_sc.shape and np.max(np.abs(w_sc - g_sc)) < TOL, f
        clear = np.abs(np.diff(w_sc, axis=1)).min(axis=1) > 2 * TOL if w_sc.shape[1] > 1 else np.ones(len(w_sc), bool)
        n_unclear = int((~clear).sum())
        all_clear &= n_unclear == 0
        for key in ("all_top_patient_ids", "all_top_labels"):
            g_rows, w_rows = got[key], want[key]
            assert all(g_rows[i] == w_rows[i] for i in np.flatnonzero(clear)), (f, key)
            assert all(sorted(g_rows[i]) == sorted(w_rows[i]) or not clear[i] for i in range(len(w_rows)))
        for key in scalar:
            if n_unclear == 0:
                assert abs(got[key] - want[key]) < 1e-12, (f, key, got[key], want[key])
            elif not key.startswith("macro"):
                assert abs(got[key] - want[key]) <= n_unclear / len(w_sc) + 1e-12, (f, key)
        if n_unclear == 0:
            assert got["confusion_matrix_top1"] == want["confusion_matrix_top1"]
            assert got["confusion_matrix_vote"] == want["confusion_matrix_vote"]
        assert set(got) - set(scalar) >= {"all_top_labels", "all_top_scores", "all_top_patient_ids", "test_patient_ids"}
    got_csv, want_csv = (exp / "summary.csv").read_text(), open(os.path.join(gold, "summary.csv")).read()
    assert got_csv.splitlines()[0] == want_csv.splitlines()[0] and len(got_csv.splitlines()) == len(want_csv.splitlines())
    if all_clear:
        assert got_csv == want_csv                                 # byte-identical summary
    print(f"{tag}: all rows clear = {all_clear}")
