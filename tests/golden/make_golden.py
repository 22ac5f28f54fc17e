"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (build container only).

    python tests/golden/make_golden.py

imports ``retrieval`` and ``utils.cv_evaluator`` from ``/root/reference`` (with
``matplotlib``/``seaborn`` stubbed -- they are only used for a PNG plot) and
writes small ``.npz`` fixtures next to this script.  The reference tree does
not exist on the GPU box; tests only read the committed fixtures.

Labels are stored as integer codes; the reference is fed ``class_<code>``
strings.  Unseeded PCA (utils/cv_evaluator.py:89) is neutralised with
``np.random.seed`` immediately before each reference call.
"""
import importlib.machinery as im
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("EMR2A_REFERENCE", "/root/reference")


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__spec__ = im.ModuleSpec(name, None)
    m.__path__ = []
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def load_reference():
    sys.dont_write_bytecode = True
    plt = _stub("matplotlib.pyplot", subplots=None, tight_layout=None, savefig=None, close=None)
    _stub("matplotlib", pyplot=plt)
    _stub("seaborn", heatmap=None)
    sys.path.insert(0, REF)
    import retrieval  # noqa
    import utils.cv_evaluator  # noqa
    import utils.metrics  # noqa
    import utils.common  # noqa
    return retrieval, sys.modules["utils.cv_evaluator"], sys.modules["utils.metrics"], sys.modules["utils.common"]


def names(codes, n_classes):
    return [f"class_{int(c)}" for c in codes]


def code_of(name):
    return int(name.split("_")[1])


def main():
    sys.path.insert(0, REPO)
    from emr2a_b200 import synth
    sys.path.remove(REPO)
    retrieval, cvmod, metrics, common = load_reference()

    # ---------------- primitives ----------------
    rng = np.random.default_rng(101)
    out = {}
    q = rng.standard_normal(40).astype(np.float32) * 3.0
    db = rng.standard_normal((57, 40)).astype(np.float32) * 0.7
    db[5] = 0.0                                   # zero row -> epsilon path
    db[9] = db[3]                                  # duplicate row -> exact tie
    out["cos_q"], out["cos_db"] = q, db
    out["cos_out"] = retrieval.compute_cosine_similarity(q, db)
    out["euc_out"] = retrieval.compute_euclidean_similarity(q, db)
    t = rng.standard_normal((33, 24)).astype(np.float32)
    i = rng.standard_normal((33, 40)).astype(np.float32) * 2.0
    t[4] = 0.0
    i[4] = 0.0
    out["ef_text"], out["ef_image"] = t, i
    out["ef_out_11"] = retrieval.early_fusion(t, i)
    out["ef_out_w"] = retrieval.early_fusion(t, i, 0.4, 0.6)
    ts = rng.standard_normal(57).astype(np.float32) * 0.1
    is_ = rng.standard_normal(57).astype(np.float32) * 0.2
    out["lf_ts"], out["lf_is"] = ts, is_
    from retrieval.fusion import normalize_scores
    for mode in ("none", "zscore", "minmax"):
        out[f"lf_out_{mode}"] = retrieval.late_fusion(ts, is_, 0.4, mode)
        out[f"ns_out_{mode}"] = normalize_scores(ts, mode)
    out["lf_out_w07"] = retrieval.late_fusion(ts, is_, 0.7)
    ev = cvmod.CVRetrievalEvaluator()
    out["nr_out"] = ev._normalize_rows(i)
    a = ev._normalize_rows(rng.standard_normal((33, 16)).astype(np.float32))
    b = ev._normalize_rows(rng.standard_normal((33, 24)).astype(np.float32))
    out["cf_img"], out["cf_txt"] = a, b
    out["cf_out"] = ev.concat_fusion(a, b)
    out["dot_out"] = ev.compute_cosine_similarity(out["cf_out"][2], out["cf_out"])
    out["l2_out"] = common.l2_normalize(q)
    out["l2_zero_out"] = common.l2_normalize(np.zeros(7, np.float32))
    out["ce_out"] = common.concat_embeddings(t[0], i[0], 0.3, 0.9)
    # retrieve_topk
    lab_codes = rng.integers(0, 3, size=33)
    tl, tsc, tid = ev.retrieve_topk(out["cf_out"][2], out["cf_out"], names(lab_codes, 3), 5)
    out["rt_labels"] = lab_codes
    out["rt_top_labels"] = np.array([code_of(x) for x in tl])
    out["rt_top_scores"] = np.array(tsc, dtype=np.float64)
    out["rt_top_idx"] = np.array([int(x.split("_")[1]) for x in tid])
    # metrics
    pred = rng.integers(0, 4, size=80)
    truth = rng.integers(0, 3, size=80)          # class 3 never true
    labs = [f"class_{c}" for c in range(4)]
    prf = metrics.compute_precision_recall_f1(names(pred, 4), names(truth, 4), labs)
    cm = metrics.compute_confusion_matrix(names(pred, 4), names(truth, 4), labs)
    out["m_pred"], out["m_truth"] = pred, truth
    out["m_prf"] = np.array([[prf[l]["precision"], prf[l]["recall"], prf[l]["f1"], prf[l]["support"]] for l in labs])
    out["m_cm"] = np.array([[cm[a_][b_] for b_ in labs] for a_ in labs])
    out["m_acc"] = metrics.compute_accuracy(names(pred, 4), names(truth, 4))
    pl = [[f"class_{c}" for c in rng.integers(0, 4, size=5)] for _ in range(80)]
    out["m_predlists"] = np.array([[code_of(x) for x in row] for row in pl])
    out["m_top3"] = metrics.compute_top_k_accuracy(pl, names(truth, 4), 3)
    # votes
    vl = rng.integers(0, 3, size=(200, 5))
    vs = np.sort(rng.random((200, 5)).astype(np.float32), axis=1)[:, ::-1]
    vs[::7, 1] = vs[::7, 0]                        # equal scores
    true = rng.integers(0, 3, size=200)
    out["v_labels"], out["v_scores"], out["v_true"] = vl, vs, true
    tlab = [names(r, 3) for r in vl]
    tsc = [[float(x) for x in r] for r in vs]
    out["v_acc_major"] = ev.compute_vote_accuracy(tlab, tsc, names(true, 3), weighted=False)
    out["v_acc_weight"] = ev.compute_vote_accuracy(tlab, tsc, names(true, 3), weighted=True)
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **out)

    # ---------------- CV evaluator (small C1-shaped) ----------------
    n, d_img, d_txt, n_cls, pca_dim, top_k = 300, 48, 40, 3, 16, 5
    data = synth.two_modal(n, d_img, d_txt, n_cls, seed=7, sep=0.35)
    ids = synth.patient_ids(n)
    labels = names(data["labels"], n_cls)
    emb = {pid: {"image": data["image"][j], "text": data["text"][j]} for j, pid in enumerate(ids)}
    cv = {"image": data["image"], "text": data["text"], "labels": data["labels"],
          "meta": np.array([n, d_img, d_txt, n_cls, pca_dim, top_k])}
    ev = cvmod.CVRetrievalEvaluator(cv_folds=5, pca_dim=pca_dim, top_k=top_k, seed=42)
    splits = ev.stratified_split(ids, labels)
    pid_idx = {p: j for j, p in enumerate(ids)}
    for f, (tr, te) in enumerate(splits):
        tr_i = np.array([pid_idx[p] for p in tr])
        te_i = np.array([pid_idx[p] for p in te])
        cv[f"f{f}_train_idx"], cv[f"f{f}_test_idx"] = tr_i, te_i
        np.random.seed(1000 + f)
        a_tr, a_te = ev.process_embeddings(data["image"][tr_i], data["image"][te_i])
        b_tr, b_te = ev.process_embeddings(data["text"][tr_i], data["text"][te_i])
        cv[f"f{f}_img_tr"], cv[f"f{f}_img_te"] = a_tr, a_te
        cv[f"f{f}_txt_tr"], cv[f"f{f}_txt_te"] = b_tr, b_te
        for fusion, w in (("concat", 0.5), ("late", 0.3), ("image_only", 0.5), ("text_only", 0.5)):
            np.random.seed(1000 + f)
            r = ev.evaluate_fold(data["image"][tr_i], data["text"][tr_i], data["image"][te_i], data["text"][te_i],
                                 [labels[j] for j in tr_i], [labels[j] for j in te_i], list(te),
                                 fusion=fusion, top_k_list=[1, 3, 5, 5], w_text=w, train_ids=list(tr))
            key = f"f{f}_{fusion}"
            tr_pos = {p: j for j, p in enumerate(tr)}
            cv[key + "_top_idx"] = np.array([[tr_pos[p] for p in row] for row in r["all_top_patient_ids"]])
            cv[key + "_top_scores"] = np.array(r["all_top_scores"], dtype=np.float64)
            cv[key + "_top_labels"] = np.array([[code_of(x) for x in row] for row in r["all_top_labels"]])
            cv[key + "_metrics"] = np.array([r["top1"], r["top3"], r["top5"], r["vote_acc"], r["weighted_vote_acc"],
                                             r["macro_precision"], r["macro_recall"], r["macro_f1"]], dtype=np.float64)
            cls = [f"class_{c}" for c in range(n_cls)]
            cv[key + "_cm_top1"] = np.array([[r["confusion_matrix_top1"][a_][b_] for b_ in cls] for a_ in cls])
            cv[key + "_cm_vote"] = np.array([[r["confusion_matrix_vote"][a_][b_] for b_ in cls] for a_ in cls])
    # whole run_cv with a single seed up front (what the drop-in run_cv is compared with)
    for fusion, w in (("concat", 0.5), ("late", 0.25)):
        np.random.seed(77)
        ev = cvmod.CVRetrievalEvaluator(cv_folds=5, pca_dim=pca_dim, top_k=top_k, seed=42)
        res = ev.run_cv(ids, labels, emb, fusion=fusion, top_k_list=[1, 3, 5, 5], w_text=w)
        cv[f"runcv_{fusion}_summary"] = np.array(
            [[res["summary"][m][s] for s in ("mean", "std", "min", "max")]
             for m in ("top1", "top3", "top5", "vote_acc", "weighted_vote_acc",
                       "macro_precision", "macro_recall", "macro_f1")])
        cv[f"runcv_{fusion}_fold_top1"] = np.array([r["top1"] for r in res["fold_results"]])
        cv[f"runcv_{fusion}_f0_scores"] = np.array(res["fold_results"][0]["all_top_scores"])
        cv[f"runcv_{fusion}_f0_ids"] = np.array(
            [[pid_idx[p] for p in row] for row in res["fold_results"][0]["all_top_patient_ids"]])
        # every fold: Top-K scores / ids and the eight metrics (the drop-in must reproduce them exactly on clear rows)
        cv[f"runcv_{fusion}_scores"] = np.array([r["all_top_scores"] for r in res["fold_results"]])
        cv[f"runcv_{fusion}_ids"] = np.array(
            [[[pid_idx[p] for p in row] for row in r["all_top_patient_ids"]] for r in res["fold_results"]])
        cv[f"runcv_{fusion}_fold_metrics"] = np.array(
            [[r[m] for m in ("top1", "top3", "top5", "vote_acc", "weighted_vote_acc",
                             "macro_precision", "macro_recall", "macro_f1")] for r in res["fold_results"]])
        if fusion == "concat":
            keys = sorted(res["fold_results"][0].keys())
            with open(os.path.join(HERE, "cv_result_keys.json"), "w") as fh:
                json.dump({"fold_keys": keys, "summary_keys": list(res["summary"].keys())}, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "cv_small.npz"), **cv)

    # ---------------- hold-out evaluator ----------------
    ho = {}
    tr = synth.two_modal(200, 32, 24, 3, seed=23, sep=0.4)
    te = synth.two_modal(60, 32, 24, 3, seed=29, sep=0.4)
    # same class centres for train and test: regenerate test from the train generator's tail
    al = synth.two_modal(260, 32, 24, 3, seed=23, sep=0.4)
    tr = {k: v[:200] for k, v in al.items()}
    te = {k: v[200:] for k, v in al.items()}
    # raw, un-normalised, different scales per row
    scale = np.random.default_rng(5).uniform(0.5, 4.0, size=(260, 1)).astype(np.float32)
    tr_img, te_img = tr["image"] * scale[:200], te["image"] * scale[200:]
    tr_txt, te_txt = tr["text"] * scale[:200], te["text"] * scale[200:]
    ho.update(tr_img=tr_img, te_img=te_img, tr_txt=tr_txt, te_txt=te_txt,
              tr_labels=tr["labels"], te_labels=te["labels"])
    rev = retrieval.RetrievalEvaluator()
    trl, tel = names(tr["labels"], 3), names(te["labels"], 3)
    runs = {
        "early": dict(fusion_type="early", text_weight=0.4),
        "late_none": dict(fusion_type="late", text_weight=0.4, score_mode="none"),
        "late_zscore": dict(fusion_type="late", text_weight=0.3, score_mode="zscore"),
        "late_minmax": dict(fusion_type="late", text_weight=0.6, score_mode="minmax"),
    }
    for name, kw in runs.items():
        r = rev.evaluate_retrieval(tr_txt, te_txt, tr_img, te_img, trl, tel, top_k_list=[1, 3, 5, 7], **kw)
        sc = {k: v for k, v in r.items() if not isinstance(v, list)}
        ho[name + "_keys"] = np.array(sorted(sc.keys()))
        ho[name + "_vals"] = np.array([sc[k] for k in sorted(sc.keys())], dtype=np.float64)
        if "all_top_labels_top5" in r:
            ho[name + "_top5"] = np.array([[code_of(x) for x in row] for row in r["all_top_labels_top5"]])
    # the score matrices behind those metrics, from the reference's own primitives (retrieval/similarity.py:4-7,
    # retrieval/fusion.py:4-28), so that the parity test can tell which queries sit on a near-tie
    cos_t = np.array([retrieval.compute_cosine_similarity(te_txt[i], tr_txt) for i in range(len(tel))])
    cos_i = np.array([retrieval.compute_cosine_similarity(te_img[i], tr_img) for i in range(len(tel))])
    ho["scores_text"], ho["scores_image"] = cos_t, cos_i
    f_db = retrieval.early_fusion(tr_txt, tr_img, 0.4, 1 - 0.4)
    f_q = retrieval.early_fusion(te_txt, te_img, 0.4, 1 - 0.4)
    ho["scores_early"] = np.array([retrieval.compute_cosine_similarity(f_q[i], f_db) for i in range(len(tel))])
    for name, kw in runs.items():
        if name != "early":
            ho["scores_" + name] = np.array([retrieval.late_fusion(cos_t[i], cos_i[i], kw["text_weight"], kw["score_mode"])
                                             for i in range(len(tel))])
    r = rev.evaluate_retrieval(None, None, tr_img, te_img, trl, tel, fusion_type="none", top_k_list=[1, 3, 5, 5])
    ho["imgonly_keys"] = np.array(sorted(r.keys()))
    ho["imgonly_vals"] = np.array([r[k] for k in sorted(r.keys())], dtype=np.float64)
    # from-scores helpers
    sc = np.random.default_rng(9).standard_normal((60, 200)).astype(np.float32)
    ho["fs_scores"] = sc
    ho["fs_top3"] = rev._compute_top_k_accuracy_from_scores(sc, trl, tel, 3)
    ho["fs_weighted"] = rev._compute_weighted_accuracy_from_scores(sc, trl, tel)
    ho["fs_top5_labels"] = np.array([[code_of(x) for x in row] for row in rev.get_all_top_labels(sc, trl, tel, 5)])
    # python-random stratified hold-out split (host logic; RNG stream must match)
    split_labels = names(np.random.default_rng(31).integers(0, 4, size=97), 4) + ["solo"]
    rs = retrieval.RetrievalEvaluator(test_ratio=0.2, seed=42)
    tr_i, te_i = rs.stratified_split(split_labels)
    tr_j, te_j = rs.stratified_split(split_labels)        # second call continues the RNG stream
    ho["ss_labels"] = np.array([4 if x == "solo" else code_of(x) for x in split_labels])
    ho["ss_train1"], ho["ss_test1"], ho["ss_train2"], ho["ss_test2"] = map(np.array, (tr_i, te_i, tr_j, te_j))
    np.savez_compressed(os.path.join(HERE, "holdout_small.npz"), **ho)

    import sklearn
    with open(os.path.join(HERE, "VERSIONS.json"), "w") as fh:
        json.dump({"numpy": np.__version__, "sklearn": sklearn.__version__,
                   "reference": "Ali-Xiyao/emr2a-evidence-grounded-multimodal-retrieval @ /root/reference"}, fh, indent=1)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
