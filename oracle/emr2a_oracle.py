"""CPU oracle for the EMR2A retrieval hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm
(Ali-Xiyao/emr2a-evidence-grounded-multimodal-retrieval).  It is the checker
for the CUDA path, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  Nothing under ``emr2a_b200/`` does.

Parity pin: the reference ships no tests and no golden vectors,
so the pin is the reference itself, executed in the build container by
``tests/golden/make_golden.py``; its inputs/outputs are committed under
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks every
function here against them.

Integer label codes replace the reference's ``List[str]`` labels: code ``c``
stands for ``sorted(set(labels))[c]``.  All comparisons the reference makes on
strings are equality tests (plus one ``sorted`` for the class list), so the
mapping is exact.

Tie rule: the reference ranks with ``np.argsort(x)[-k:][::-1]`` (unstable
introsort; order among exactly equal scores is implementation defined).  The
oracle fixes the order to (score descending, database index ascending), which
equals the reference wherever the top-k scores and the (k+1)-th are distinct.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

EPS = 1e-8  # the reference's additive epsilon (python float; a no-op in fp32 for norms >~ 0.1)


# --------------------------------------------------------------------------
# row normalisation / fusion
# --------------------------------------------------------------------------
def unit_rows(mat: np.ndarray) -> np.ndarray:
    """``arr / (||row||_2 + 1e-8)``.

    Follows utils/cv_evaluator.py:95-97 and retrieval/evaluator.py:75-77.
    """
    lengths = np.linalg.norm(mat, axis=1, keepdims=True) + EPS
    return mat / lengths


def fuse_concat_cv(img: np.ndarray, txt: np.ndarray) -> np.ndarray:
    """Image-first concatenation followed by row normalisation.

    Follows utils/cv_evaluator.py:99-105.
    """
    return unit_rows(np.concatenate([img, txt], axis=1))


def fuse_early(text: np.ndarray, image: np.ndarray,
               text_weight: float = 1.0, image_weight: float = 1.0) -> np.ndarray:
    """Text-first weighted concatenation followed by row normalisation.

    Follows retrieval/fusion.py:17-28.
    """
    joined = np.concatenate([text * text_weight, image * image_weight], axis=-1)
    return joined / (np.linalg.norm(joined, axis=1, keepdims=True) + EPS)


def rescale_scores(scores: np.ndarray, mode: str = "none") -> np.ndarray:
    """Per-query score normalisation (none / zscore / minmax).

    Follows retrieval/fusion.py:31-42: moments are taken with numpy in the
    array dtype, widened to python float, and applied back in the array dtype.
    """
    if mode == "zscore":
        mu = float(scores.mean())
        sd = float(scores.std())
        return (scores - mu) / (sd + EPS)
    if mode == "minmax":
        lo = float(scores.min())
        hi = float(scores.max())
        return (scores - lo) / (hi - lo + EPS)
    return scores


def fuse_late_scores(text_scores: np.ndarray, image_scores: np.ndarray,
                     text_weight: float = 0.4, score_mode: str = "none") -> np.ndarray:
    """``w * norm(text) + (1 - w) * norm(image)``.  Follows retrieval/fusion.py:4-14."""
    t = rescale_scores(text_scores, score_mode)
    i = rescale_scores(image_scores, score_mode)
    return text_weight * t + (1 - text_weight) * i


def unit_vector(vec: np.ndarray) -> np.ndarray:
    """Single-vector normalise with a zero guard and no epsilon (utils/common.py:4-8)."""
    length = np.linalg.norm(vec)
    return vec if length == 0 else vec / length


def fuse_single(text_vec: np.ndarray, image_vec: np.ndarray,
                text_weight: float = 1.0, image_weight: float = 1.0) -> np.ndarray:
    """Text-first weighted concat of two vectors, then unit_vector (utils/common.py:11-22)."""
    joined = np.concatenate([text_vec * float(text_weight), image_vec * float(image_weight)], axis=0)
    return unit_vector(joined)


# --------------------------------------------------------------------------
# similarity
# --------------------------------------------------------------------------
def cosine_one_vs_db(query: np.ndarray, database: np.ndarray) -> np.ndarray:
    """Cosine of one query against every database row, both re-normalised
    with the additive epsilon.  Follows retrieval/similarity.py:4-7."""
    q = query / (np.linalg.norm(query) + EPS)
    d = database / (np.linalg.norm(database, axis=1, keepdims=True) + EPS)
    return np.dot(d, q)


def euclid_one_vs_db(query: np.ndarray, database: np.ndarray) -> np.ndarray:
    """``1 - dist / max(dist)``.  Follows retrieval/similarity.py:10-15."""
    dist = np.linalg.norm(database - query, axis=1)
    far = np.max(dist)
    return 1.0 - dist / far if far > 0 else 1.0 - dist


def dot_one_vs_db(query: np.ndarray, database: np.ndarray) -> np.ndarray:
    """Plain ``db @ q`` on pre-normalised rows (utils/cv_evaluator.py:107-112)."""
    return np.dot(database, query)


# --------------------------------------------------------------------------
# Top-K and votes
# --------------------------------------------------------------------------
def topk_desc(scores: np.ndarray, k: int) -> np.ndarray:
    """Indices of the ``k`` best scores, best first.

    The reference is ``np.argsort(scores)[-k:][::-1]`` (utils/cv_evaluator.py:123,
    :237; retrieval/evaluator.py:189,204,220,243,272).  Ties are broken here by
    ascending index (see module docstring).  ``k > len(scores)`` returns all.
    """
    order = np.argsort(-scores, kind="stable")
    return order[:k]


def vote_majority(labels: Sequence[int]) -> int:
    """``Counter(labels).most_common(1)[0][0]``: highest count, ties go to the
    label whose first occurrence has the best rank (utils/cv_evaluator.py:150,
    :252, :285)."""
    seen: Dict[int, int] = {}
    for lab in labels:
        seen[lab] = seen.get(lab, 0) + 1
    best, best_n = None, -1
    for lab, n in seen.items():       # dict keeps first-occurrence order
        if n > best_n:
            best, best_n = lab, n
    return best


def vote_weighted(labels: Sequence[int], scores: Sequence, acc: str = "f64") -> int:
    """Per-label score sums in rank order; the largest sum wins, ties go to the
    label inserted first.

    ``acc='f64'``: scores went through ``float()`` first (utils/cv_evaluator.py:125,
    :239, :142-147, :255-260, :288-293).  ``acc='f32'``: numpy float32 scalars
    are summed as they are (retrieval/evaluator.py:224-230, :247-253).
    """
    sums: Dict[int, object] = {}
    for lab, sc in zip(labels, scores):
        sc = float(sc) if acc == "f64" else np.float32(sc)
        if lab not in sums:
            sums[lab] = 0.0
        sums[lab] = sums[lab] + sc
    best, best_s = None, None
    for lab, s in sums.items():
        if best_s is None or s > best_s:
            best, best_s = lab, s
    return best


# --------------------------------------------------------------------------
# metrics (utils/metrics.py)
# --------------------------------------------------------------------------
def confusion_counts(pred: Sequence[int], truth: Sequence[int], n_classes: int) -> np.ndarray:
    """``matrix[true, pred] += 1`` (utils/metrics.py:56-75)."""
    m = np.zeros((n_classes, n_classes), dtype=np.int64)
    for p, t in zip(pred, truth):
        if 0 <= p < n_classes and 0 <= t < n_classes:
            m[t, p] += 1
    return m


def prf_per_class(pred: Sequence[int], truth: Sequence[int], n_classes: int) -> List[Dict[str, float]]:
    """Per-class precision / recall / F1 / support with 0.0 on empty
    denominators (utils/metrics.py:30-53)."""
    pred = np.asarray(pred)
    truth = np.asarray(truth)
    out = []
    for c in range(n_classes):
        tp = int(np.sum((pred == c) & (truth == c)))
        fp = int(np.sum((pred == c) & (truth != c)))
        fn = int(np.sum((pred != c) & (truth == c)))
        p = tp / (tp + fp) if (tp + fp) > 0 else 0.0
        r = tp / (tp + fn) if (tp + fn) > 0 else 0.0
        f = 2 * p * r / (p + r) if (p + r) > 0 else 0.0
        out.append({"precision": p, "recall": r, "f1": f, "support": int(np.sum(truth == c))})
    return out


# --------------------------------------------------------------------------
# evaluators
# --------------------------------------------------------------------------
def cv_fold_eval(
    db_img: Optional[np.ndarray], db_txt: Optional[np.ndarray],
    q_img: Optional[np.ndarray], q_txt: Optional[np.ndarray],
    db_labels: np.ndarray, q_labels: np.ndarray, n_classes: int,
    fusion: str = "concat", top_k: int = 5, top_k_list: Sequence[int] = (1, 3, 5, 5),
    w_text: float = 0.5,
) -> Dict:
    """One CV fold from the *processed* (post StandardScaler/PCA/row-normalise)
    arrays onward.  Follows utils/cv_evaluator.py:186-334.

    Returns integer/array results; the drop-in layer maps them back to the
    reference's string/list structures.
    """
    if fusion == "image_only":
        db, qs = db_img, q_img
    elif fusion == "text_only":
        db, qs = db_txt, q_txt
    elif fusion == "concat":
        db, qs = fuse_concat_cv(db_img, db_txt), fuse_concat_cv(q_img, q_txt)
    elif fusion == "late":
        db = qs = None
    else:
        raise ValueError(f"Unknown fusion type: {fusion}")

    n_q = len(q_labels)
    kk = min(top_k, len(db_labels))
    top_idx = np.zeros((n_q, kk), dtype=np.int64)
    top_scores = np.zeros((n_q, kk), dtype=np.float64)
    pred_top1 = np.zeros(n_q, dtype=np.int64)
    pred_vote = np.zeros(n_q, dtype=np.int64)
    pred_wvote = np.zeros(n_q, dtype=np.int64)
    hits = {int(k): np.zeros(n_q, dtype=np.int64) for k in top_k_list}

    for i in range(n_q):
        if fusion == "late":
            s_img = dot_one_vs_db(q_img[i], db_img)              # :233
            s_txt = dot_one_vs_db(q_txt[i], db_txt)              # :234
            sims = w_text * s_txt + (1 - w_text) * s_img         # :235
        else:
            sims = dot_one_vs_db(qs[i], db)                      # :122
        idx = topk_desc(sims, top_k)                             # :123 / :237
        labs = [int(db_labels[j]) for j in idx]
        scs = [float(sims[j]) for j in idx]
        top_idx[i, :len(idx)] = idx
        top_scores[i, :len(idx)] = scs
        pred_top1[i] = labs[0]                                   # :249 / :282
        pred_vote[i] = vote_majority(labs)                       # :252 / :285
        pred_wvote[i] = vote_weighted(labs, scs, "f64")          # :255-260 / :288-293
        for k in hits:
            hits[k][i] = 1 if int(q_labels[i]) in labs[:k] else 0   # :263-267 / :296-300

    q_labels = np.asarray(q_labels)
    res: Dict = {f"top{k}": float(np.mean(v)) for k, v in hits.items()}
    res["vote_acc"] = float(np.mean(pred_vote == q_labels))          # :305-307
    res["weighted_vote_acc"] = float(np.mean(pred_wvote == q_labels))  # :308-310
    prf = prf_per_class(pred_vote, q_labels, n_classes)              # :314-316
    res["macro_precision"] = float(np.mean([v["precision"] for v in prf]))
    res["macro_recall"] = float(np.mean([v["recall"] for v in prf]))
    res["macro_f1"] = float(np.mean([v["f1"] for v in prf]))
    res["confusion_top1"] = confusion_counts(pred_top1, q_labels, n_classes)   # :322-324
    res["confusion_vote"] = confusion_counts(pred_vote, q_labels, n_classes)   # :325-327
    res.update(top_idx=top_idx, top_scores=top_scores, pred_top1=pred_top1,
               pred_vote=pred_vote, pred_weighted=pred_wvote,
               hits={k: v.copy() for k, v in hits.items()})
    return res


def _holdout_topk_acc(db: np.ndarray, qs: np.ndarray, db_labels, q_labels, k: int) -> float:
    """retrieval/evaluator.py:178-193."""
    good = 0
    for i in range(len(qs)):
        sims = cosine_one_vs_db(qs[i], db)
        labs = [int(db_labels[j]) for j in topk_desc(sims, k)]
        good += int(int(q_labels[i]) in labs)
    return good / len(q_labels)


def _holdout_weighted_acc(db: np.ndarray, qs: np.ndarray, db_labels, q_labels) -> float:
    """retrieval/evaluator.py:210-233 (K fixed at 5, fp32 sums)."""
    good = 0
    for i in range(len(qs)):
        sims = cosine_one_vs_db(qs[i], db)
        idx = topk_desc(sims, 5)
        labs = [int(db_labels[j]) for j in idx]
        good += int(vote_weighted(labs, [sims[j] for j in idx], "f32") == int(q_labels[i]))
    return good / len(q_labels)


def _scores_topk_acc(scores: np.ndarray, db_labels, q_labels, k: int) -> float:
    """retrieval/evaluator.py:195-208."""
    good = 0
    for i in range(len(scores)):
        labs = [int(db_labels[j]) for j in topk_desc(scores[i], k)]
        good += int(int(q_labels[i]) in labs)
    return good / len(q_labels)


def _scores_weighted_acc(scores: np.ndarray, db_labels, q_labels) -> float:
    """retrieval/evaluator.py:235-256."""
    good = 0
    for i in range(len(scores)):
        idx = topk_desc(scores[i], 5)
        labs = [int(db_labels[j]) for j in idx]
        good += int(vote_weighted(labs, [scores[i][j] for j in idx], "f32") == int(q_labels[i]))
    return good / len(q_labels)


def holdout_eval(
    db_txt: Optional[np.ndarray], q_txt: Optional[np.ndarray],
    db_img: Optional[np.ndarray], q_img: Optional[np.ndarray],
    db_labels, q_labels, text_weight: float = 0.4, fusion_type: str = "late",
    score_mode: str = "none", top_k_list: Sequence[int] = (1, 3, 5),
) -> Dict:
    """Hold-out evaluator.  Follows retrieval/evaluator.py:94-176; the
    ``all_top_labels_top5`` entry holds label codes."""
    out: Dict = {}
    if fusion_type == "early":
        if db_txt is None or q_txt is None or db_img is None or q_img is None:
            raise ValueError("Early fusion requires both text and image embeddings")
        fdb = fuse_early(db_txt, db_img, text_weight, 1 - text_weight)
        fq = fuse_early(q_txt, q_img, text_weight, 1 - text_weight)
        for k in top_k_list:
            out[f"top{k}"] = _holdout_topk_acc(fdb, fq, db_labels, q_labels, k)
        out["weighted"] = _holdout_weighted_acc(fdb, fq, db_labels, q_labels)
        return out
    if q_txt is not None and db_txt is not None:
        for k in top_k_list:
            out[f"text_top{k}"] = _holdout_topk_acc(db_txt, q_txt, db_labels, q_labels, k)
        out["text_weighted"] = _holdout_weighted_acc(db_txt, q_txt, db_labels, q_labels)
    if q_img is not None and db_img is not None:
        for k in top_k_list:
            out[f"image_top{k}"] = _holdout_topk_acc(db_img, q_img, db_labels, q_labels, k)
        out["image_weighted"] = _holdout_weighted_acc(db_img, q_img, db_labels, q_labels)
    if q_txt is not None and q_img is not None:
        rows = []
        for i in range(len(q_labels)):
            st = cosine_one_vs_db(q_txt[i], db_txt)
            si = cosine_one_vs_db(q_img[i], db_img)
            rows.append(fuse_late_scores(st, si, text_weight, score_mode))
        fused = np.array(rows)
        for k in top_k_list:
            out[f"top{k}"] = _scores_topk_acc(fused, db_labels, q_labels, k)
        out["weighted"] = _scores_weighted_acc(fused, db_labels, q_labels)
        out["all_top_labels_top5"] = [
            [int(db_labels[j]) for j in topk_desc(fused[i], 5)] for i in range(len(fused))]
        out["_fused_scores"] = fused
    return out


# --------------------------------------------------------------------------
# batched helpers for parity tests / CPU-baseline timing
# --------------------------------------------------------------------------
def search_topk_batched(qs: np.ndarray, db: np.ndarray, k: int,
                        q_fold: Optional[np.ndarray] = None,
                        db_fold: Optional[np.ndarray] = None,
                        block: int = 256) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k of ``db @ q`` for many queries at once (sgemm instead of the
    reference's per-query sgemv: same math, summation order differs by
    <= ~1e-7).  Rows with ``db_fold == q_fold`` are excluded (the CV rule
    "never retrieve from the query's own fold", utils/cv_evaluator.py:349-376).
    Returns (indices int64 [Q,k], scores float32 [Q,k]); -1 pads short rows."""
    n_q = qs.shape[0]
    idx_out = np.full((n_q, k), -1, dtype=np.int64)
    sc_out = np.zeros((n_q, k), dtype=np.float32)
    for s in range(0, n_q, block):
        sims = qs[s:s + block] @ db.T
        if q_fold is not None:
            sims = np.where(q_fold[s:s + block, None] == db_fold[None, :], -np.inf, sims)
        for r in range(sims.shape[0]):
            row = sims[r]
            if k < row.shape[0]:
                cand = np.argpartition(-row, k)[:k + 1]
                # widen to every index tied with the k-th value so the tie rule is global
                kth = np.sort(row[cand])[::-1][k - 1]
                cand = np.nonzero(row >= kth)[0]
            else:
                cand = np.arange(row.shape[0])
            order = cand[np.argsort(-row[cand], kind="stable")][:k]
            order = order[np.isfinite(row[order])]
            idx_out[s + r, :len(order)] = order
            sc_out[s + r, :len(order)] = row[order]
    return idx_out, sc_out


def reference_style_search_and_vote(qs: np.ndarray, db: np.ndarray, db_labels: np.ndarray,
                                    q_labels: np.ndarray, k: int) -> Dict:
    """The per-query loop exactly as the reference runs it (sgemv + full argsort
    + python votes): this is what ``bench.py`` times as the CPU baseline.
    Follows utils/cv_evaluator.py:269-300."""
    n_q = qs.shape[0]
    top1 = np.zeros(n_q, np.int64)
    vote = np.zeros(n_q, np.int64)
    wvote = np.zeros(n_q, np.int64)
    idx_all = np.zeros((n_q, k), np.int64)
    for i in range(n_q):
        sims = np.dot(db, qs[i])
        idx = np.argsort(sims)[-k:][::-1]
        labs = [int(db_labels[j]) for j in idx]
        scs = [float(sims[j]) for j in idx]
        idx_all[i] = idx
        top1[i] = labs[0]
        vote[i] = vote_majority(labs)
        wvote[i] = vote_weighted(labs, scs, "f64")
    return {"top_idx": idx_all, "pred_top1": top1, "pred_vote": vote, "pred_weighted": wvote,
            "top1": float(np.mean(top1 == q_labels)), "vote_acc": float(np.mean(vote == q_labels)),
            "weighted_vote_acc": float(np.mean(wvote == q_labels))}


# --------------------------------------------------------------------------
# per-fold preprocessing: StandardScaler -> PCA -> row normalisation
# --------------------------------------------------------------------------
# The arithmetic lives in scikit-learn (un-vendored, unpinned by the reference's
# requirements.txt:1-18; this container has 1.9.0, tests/golden/VERSIONS.json).  The restatement
# below follows sklearn's published algorithm:
#   StandardScaler  float64 column mean and population variance (_incremental_mean_and_var reduces float32 input
#                   in float64), near-constant features get scale 1 (_is_constant_feature), and transform works
#                   in the dtype of X: `X -= astype(mean_, X.dtype); X /= astype(scale_, X.dtype)`;
#   PCA             principal axes of the centred data ordered by decreasing variance, sign fixed by
#                   svd_flip(u_based_decision=False); transform = X @ components^T - mean @ components^T.
# The basis is computed in float64 (covariance + symmetric eigen-decomposition), i.e. the exact answer that
# sklearn's "full" and "covariance_eigh" solvers approximate in fp32 and its unseeded "randomized" solver
# (picked by "auto" for mid-sized folds, utils/cv_evaluator.py:89) approximates differently on every run.
# Pinned by tests/test_oracle_golden.py against sklearn itself and against the reference's process_embeddings
# outputs stored in tests/golden/cv_small.npz (240 x 48 / 240 x 40 folds -> sklearn's "full" solver).
def scaler_fit(train: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(mean_, scale_) of StandardScaler.fit, float64 (utils/cv_evaluator.py:78-79)."""
    x = np.asarray(train, dtype=np.float64)
    n = x.shape[0]
    mean = x.sum(axis=0) / n
    dev = x - mean
    var = ((dev ** 2).sum(axis=0) - dev.sum(axis=0) ** 2 / n) / n
    eps = np.finfo(np.float64).eps
    constant = var <= n * eps * var + (n * mean * eps) ** 2
    scale = np.sqrt(var)
    scale[constant] = 1.0
    return mean, scale


def scaler_apply(x: np.ndarray, mean: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """StandardScaler.transform in the dtype of ``x`` (utils/cv_evaluator.py:79-80)."""
    x = np.array(x, copy=True)
    x -= mean.astype(x.dtype)
    x /= scale.astype(x.dtype)
    return x


def pca_exact_fit(z_train: np.ndarray, n_components: int) -> Tuple[np.ndarray, np.ndarray]:
    """(components_ [P, D], mean_ [D]) of the exact PCA of the rows of ``z_train`` (utils/cv_evaluator.py:89-90)."""
    z = np.asarray(z_train, dtype=np.float64)
    n = z.shape[0]
    mean = z.sum(axis=0) / n
    cov = (z.T @ z - n * np.outer(mean, mean)) / (n - 1)
    cov = 0.5 * (cov + cov.T)
    _, vecs = np.linalg.eigh(cov)
    comps = vecs[:, ::-1].T[:n_components].copy()
    pivot = np.argmax(np.abs(comps), axis=1)
    comps *= np.sign(comps[np.arange(comps.shape[0]), pivot])[:, None]
    return comps, mean


def process_embeddings_exact(train: np.ndarray, test: np.ndarray, pca_dim: Optional[int]
                             ) -> Tuple[np.ndarray, np.ndarray]:
    """CVRetrievalEvaluator.process_embeddings (utils/cv_evaluator.py:73-93) with the exact PCA basis;
    ``pca_dim=None`` skips the PCA (RetrievalEvaluator.process_embeddings with use_pca=False,
    retrieval/evaluator.py:50-73)."""
    mean, scale = scaler_fit(train)
    tr = scaler_apply(np.asarray(train, dtype=np.float32), mean, scale)
    te = scaler_apply(np.asarray(test, dtype=np.float32), mean, scale)
    n_comp = min(pca_dim, tr.shape[0] - 1, tr.shape[1]) if pca_dim is not None else 0
    if n_comp > 0:
        comps, zmean = pca_exact_fit(tr, n_comp)
        w = comps.astype(np.float32)
        bias = zmean.astype(np.float32)[None, :] @ w.T
        tr = tr @ w.T - bias
        te = te @ w.T - bias
    return unit_rows(tr), unit_rows(te)


# --------------------------------------------------------------------------
# full-size sample parity: the reference loop for a SAMPLE of queries against a
# database that only exists in row chunks (C2-C5 of BASELINE.json: 1M-10M rows)
# --------------------------------------------------------------------------
class StreamedReferenceSample:
    """The reference's per-query loop for a sample of queries against a database too large to keep on the host:
    the database is fed in row chunks (``add_chunk``), each chunk is normalised and fused the way the reference
    does it for the whole matrix (row-wise operations, so chunking changes nothing), every sample query is scored
    against the chunk with one ``np.dot`` sgemv, and the scores land in a ``[sample, N]`` float32 matrix.
    ``finish`` then runs the reference's ``np.argsort(scores)[-k:][::-1]`` over all N scores of each query,
    followed by the python votes.

    mode ``concat``: utils/cv_evaluator.py:99-105 + :122-123 (:269-300 is the loop);
    mode ``late``:   utils/cv_evaluator.py:232-237 (``w*txt + (1-w)*img`` of two sgemv results, merge-then-Top-K);
    mode ``single``: one modality (:186-195).
    ``q_fold`` / chunk ``fold``: the CV rule of :349-376 -- a query is only scored against the rows of the
    OTHER folds (the reference builds per-fold train matrices; here the own-fold rows get -inf).
    Timings: ``prep_seconds`` (normalise + fuse of the database chunks), ``query_seconds`` (sgemv + argsort +
    votes, all sample queries)."""

    def __init__(self, mode: str, k: int, n_rows: int, q_img: np.ndarray, q_txt: Optional[np.ndarray] = None,
                 w_text: float = 0.5, q_fold: Optional[np.ndarray] = None):
        import time
        self._clock = time.perf_counter
        if mode not in ("concat", "late", "single"):
            raise ValueError(f"unknown mode {mode!r}")
        self.mode, self.k, self.n_rows, self.w_text = mode, int(k), int(n_rows), w_text
        self.q_fold = None if q_fold is None else np.asarray(q_fold)
        self.q = self._prepare(np.asarray(q_img, dtype=np.float32),
                               None if q_txt is None else np.asarray(q_txt, dtype=np.float32))
        self.n_q = (self.q[0] if isinstance(self.q, tuple) else self.q).shape[0]
        self.scores = np.full((self.n_q, self.n_rows), -np.inf, dtype=np.float32)
        self.prep_seconds = 0.0
        self.query_seconds = 0.0
        self.rows_seen = 0
        self.rows_scored = 0          # (query, row) pairs actually scored
        self.rows_timed = 0           # rows / pairs fed with timed=True (the bounded timing sample)
        self.pairs_timed = 0
        import threading
        self._lock = threading.Lock()

    def _prepare(self, img: np.ndarray, txt: Optional[np.ndarray]):
        if self.mode == "concat":
            return fuse_concat_cv(unit_rows(img), unit_rows(txt))
        if self.mode == "late":
            return unit_rows(img), unit_rows(txt)
        return unit_rows(img)

    def add_chunk(self, row0: int, img: np.ndarray, txt: Optional[np.ndarray] = None,
                  fold: Optional[np.ndarray] = None, timed: bool = True) -> None:
        """Feed database rows ``[row0, row0 + len(img))``.  ``timed=True``: the reference's way (one sgemv per query),
        counted in the timings.  ``timed=False``: the same scores from one sgemm for all sample queries (summation
        order differs by <= ~1e-7), not timed -- for the rows beyond the bounded timing sample; thread-safe, chunks
        may be fed from several threads."""
        t0 = self._clock()
        db = self._prepare(np.asarray(img, dtype=np.float32), None if txt is None else np.asarray(txt, dtype=np.float32))
        t1 = self._clock()
        n = (db[0] if isinstance(db, tuple) else db).shape[0]
        pieces = [(0, n, None)]
        if fold is not None:
            # contiguous runs of equal fold id (rows come fold-sorted or in a few runs); a query skips its own fold's runs
            fold = np.asarray(fold)
            cuts = np.flatnonzero(np.diff(fold.astype(np.int64))) + 1
            edges = np.concatenate([[0], cuts, [n]])
            pieces = [(int(a), int(b), int(fold[a])) for a, b in zip(edges[:-1], edges[1:])]
        scored = 0
        if timed:
            for i in range(self.n_q):
                for a, b, f in pieces:
                    if f is not None and self.q_fold is not None and f == int(self.q_fold[i]):
                        continue
                    if self.mode == "late":
                        s_img = np.dot(db[0][a:b], self.q[0][i])
                        s_txt = np.dot(db[1][a:b], self.q[1][i])
                        sims = self.w_text * s_txt + (1 - self.w_text) * s_img
                    else:
                        sims = np.dot(db[a:b], self.q[i])
                    self.scores[i, row0 + a:row0 + b] = sims
                    scored += b - a
        else:
            for a, b, f in pieces:
                if self.mode == "late":
                    sims = self.w_text * (self.q[1] @ db[1][a:b].T) + (1 - self.w_text) * (self.q[0] @ db[0][a:b].T)
                else:
                    sims = self.q @ db[a:b].T
                if f is not None and self.q_fold is not None:
                    own = self.q_fold == f
                    sims[own] = -np.inf
                    scored += (b - a) * int((~own).sum())
                else:
                    scored += (b - a) * self.n_q
                self.scores[:, row0 + a:row0 + b] = sims
        t2 = self._clock()
        with self._lock:
            if timed:
                self.prep_seconds += t1 - t0
                self.query_seconds += t2 - t1
                self.rows_timed += n
                self.pairs_timed += scored
            self.rows_seen += n
            self.rows_scored += scored

    def finish(self, db_labels: np.ndarray, q_labels: Optional[np.ndarray] = None) -> Dict:
        if self.rows_seen != self.n_rows:
            raise ValueError(f"only {self.rows_seen} of {self.n_rows} database rows were fed")
        k = self.k
        top_idx = np.zeros((self.n_q, k), np.int64)
        top_sc = np.zeros((self.n_q, k), np.float32)
        nxt = np.full(self.n_q, -np.inf, np.float32)
        top1 = np.zeros(self.n_q, np.int64)
        vote = np.zeros(self.n_q, np.int64)
        wvote = np.zeros(self.n_q, np.int64)
        top_lab = np.zeros((self.n_q, k), np.int64)
        t0 = self._clock()
        ref_rows = []
        for i in range(self.n_q):
            ref_rows.append(np.argsort(self.scores[i])[-(k + 1):][::-1])      # the reference's ranking op (+ the runner-up)
        rank_seconds = self._clock() - t0
        for i in range(self.n_q):
            s = self.scores[i]
            ref = ref_rows[i]
            # tie rule of this oracle (score descending, index ascending) over everything tied with the k-th score
            cand = np.flatnonzero(s >= s[ref[k - 1]])
            order = cand[np.argsort(-s[cand], kind="stable")][:k]
            top_idx[i], top_sc[i] = order, s[order]
            rest = s[ref[k]] if len(ref) > k else -np.inf
            if len(cand) > k:
                rest = s[ref[k - 1]]
            nxt[i] = rest
            t1 = self._clock()
            labs = [int(db_labels[j]) for j in order]
            scs = [float(s[j]) for j in order]
            top1[i] = labs[0]
            vote[i] = vote_majority(labs)
            wvote[i] = vote_weighted(labs, scs, "f64")
            rank_seconds += self._clock() - t1
            top_lab[i] = labs
        # timings: normalise/fuse and sgemv are linear in the rows, so the bounded timed sample scales to the database;
        # argsort + votes were run (and timed) on all N scores of every sample query
        prep_full = self.prep_seconds * self.n_rows / self.rows_timed if self.rows_timed else float("nan")
        sgemv_full = self.query_seconds * self.rows_scored / self.pairs_timed if self.pairs_timed else float("nan")
        out = {"top_idx": top_idx, "top_scores": top_sc, "next_score": nxt, "top_labels": top_lab, "pred_top1": top1, "pred_vote": vote,
               "pred_weighted": wvote, "prep_seconds": prep_full, "query_seconds": sgemv_full + rank_seconds,
               "seconds_per_query": (sgemv_full + rank_seconds) / max(self.n_q, 1), "rows_scored": self.rows_scored,
               "timed_rows": self.rows_timed, "timed_prep_seconds": self.prep_seconds,
               "timed_sgemv_seconds": self.query_seconds, "rank_seconds": rank_seconds}
        if q_labels is not None:
            q_labels = np.asarray(q_labels)
            out.update(top1=float(np.mean(top1 == q_labels)), vote_acc=float(np.mean(vote == q_labels)),
                       weighted_vote_acc=float(np.mean(wvote == q_labels)))
        return out


def sample_parity(ref: Dict, got_idx: np.ndarray, got_scores: np.ndarray, got_vote: Optional[np.ndarray] = None,
                  got_weighted: Optional[np.ndarray] = None, tol: float = 1e-5) -> Dict:
    """Compare a search result with ``StreamedReferenceSample.finish`` (or any dict with ``top_idx``,
    ``top_scores`` and optionally ``next_score``) under the stated bar: scores within ``tol``; index rows, majority
    vote and weighted vote identical wherever every adjacent gap among the k best scores and the runner-up
    exceeds ``2 * tol``.  Returns counts; ``ok`` is the conjunction."""
    ref_idx, ref_sc = np.asarray(ref["top_idx"]), np.asarray(ref["top_scores"], dtype=np.float64)
    got_idx, got_scores = np.asarray(got_idx), np.asarray(got_scores, dtype=np.float64)
    ladder = ref_sc
    if "next_score" in ref:
        ladder = np.concatenate([ref_sc, np.asarray(ref["next_score"], dtype=np.float64)[:, None]], axis=1)
    with np.errstate(invalid="ignore"):
        gaps = -np.diff(ladder, axis=1)
    clear = np.all((gaps > 2 * tol) | ~np.isfinite(gaps), axis=1)
    # score error: position-wise on clear rows, as sorted multisets otherwise (ties may swap neighbours)
    err = float(np.max(np.abs(np.sort(got_scores, axis=1) - np.sort(ref_sc, axis=1)))) if len(ref_sc) else 0.0
    rows_same = np.all(got_idx == ref_idx, axis=1)
    out = {"queries": int(len(ref_idx)), "clear_rows": int(clear.sum()), "max_score_err": err,
           "topk_rows_identical": float(np.mean(rows_same)) if len(rows_same) else 1.0,
           "clear_rows_identical": bool(np.all(rows_same[clear])),
           # as SETS: rows that are not clear may only differ by swaps / replacements within the tolerance
           "unclear_rows_within_tol": bool(all(
               np.max(np.abs(np.sort(got_scores[i]) - np.sort(ref_sc[i]))) <= 2 * tol for i in np.flatnonzero(~clear)))}
    ok = err <= tol and out["clear_rows_identical"] and out["unclear_rows_within_tol"]
    if got_vote is not None:
        same = np.asarray(got_vote) == np.asarray(ref["pred_vote"])
        out["vote_identical"] = float(np.mean(same))
        ok = ok and bool(np.all(same[clear]))
    if got_weighted is not None:
        same = np.asarray(got_weighted) == np.asarray(ref["pred_weighted"])
        out["weighted_vote_identical"] = float(np.mean(same))
        # identical rows carry scores that may differ by tol each, so two label sums closer than 2*k*tol may swap:
        # the weighted vote must be identical on clear rows whose two best label sums are further apart than that
        wclear = clear.copy()
        if "top_labels" in ref:
            labs = np.asarray(ref["top_labels"])
            for i in np.flatnonzero(clear):
                sums = sorted((float(ref_sc[i][labs[i] == c].sum()) for c in set(labs[i].tolist())), reverse=True)
                if len(sums) > 1 and sums[0] - sums[1] <= 2 * ref_sc.shape[1] * tol:
                    wclear[i] = False
        out["weighted_clear_rows"] = int(wclear.sum())
        ok = ok and bool(np.all(same[wclear]))
    out["ok"] = bool(ok)
    return out
