"""``utils.metrics`` (reference: utils/metrics.py:6-75): list-of-string label utilities.

These are host-side list processing, not arithmetic of the hot path; inside
``CVRetrievalEvaluator.evaluate_fold`` the same numbers are derived from the
confusion counts the K4 kernel accumulates on the GPU (``prf_from_confusion``).
Return structures and zero-division behaviour match the reference.
"""
from typing import Dict, List

import numpy as np


def _same_length(predictions, ground_truth):
    if len(predictions) != len(ground_truth):
        raise ValueError("Predictions and ground truth must have the same length")


def compute_accuracy(predictions: List[str], ground_truth: List[str]) -> float:
    _same_length(predictions, ground_truth)
    hits = int(np.count_nonzero(np.asarray(predictions, dtype=object) == np.asarray(ground_truth, dtype=object)))
    return hits / len(ground_truth)


def compute_top_k_accuracy(predictions: List[List[str]], ground_truth: List[str], k: int) -> float:
    _same_length(predictions, ground_truth)
    hits = sum(1 for ranked, truth in zip(predictions, ground_truth) if truth in ranked[:k])
    return hits / len(ground_truth)


def _codes(items, lut):
    return np.fromiter((lut.get(x, -1) for x in items), dtype=np.int64, count=len(items))


def prf_from_counts(tp: int, fp: int, fn: int, support: int) -> Dict[str, float]:
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0.0
    return {"precision": precision, "recall": recall, "f1": f1, "support": support}


def prf_from_confusion(confusion: np.ndarray, labels: List[str]) -> Dict[str, Dict[str, float]]:
    """Per-class precision/recall/F1 from a [true, pred] count matrix whose rows/cols are
    ``labels`` (every prediction and truth is one of ``labels``)."""
    out = {}
    for c, name in enumerate(labels):
        tp = int(confusion[c, c])
        out[name] = prf_from_counts(tp, int(confusion[:, c].sum()) - tp, int(confusion[c, :].sum()) - tp,
                                    int(confusion[c, :].sum()))
    return out


def compute_precision_recall_f1(predictions: List[str], ground_truth: List[str], labels: List[str]
                                ) -> Dict[str, Dict[str, float]]:
    lut = {name: i for i, name in enumerate(labels)}
    pred, truth = _codes(predictions, lut), _codes(ground_truth, lut)
    n = min(len(pred), len(truth))                     # zip() semantics of the reference
    pred_z, truth_z = pred[:n], truth[:n]
    out = {}
    for name in labels:
        c = lut[name]
        tp = int(np.count_nonzero((pred_z == c) & (truth_z == c)))
        fp = int(np.count_nonzero((pred_z == c) & (truth_z != c)))
        fn = int(np.count_nonzero((pred_z != c) & (truth_z == c)))
        out[name] = prf_from_counts(tp, fp, fn, int(np.count_nonzero(truth == c)))
    return out


def confusion_dict(matrix: np.ndarray, labels: List[str]) -> Dict[str, Dict[str, int]]:
    return {t: {p: int(matrix[i, j]) for j, p in enumerate(labels)} for i, t in enumerate(labels)}


def compute_confusion_matrix(predictions: List[str], ground_truth: List[str], labels: List[str]
                             ) -> Dict[str, Dict[str, int]]:
    lut = {name: i for i, name in enumerate(labels)}
    n = len(labels)
    pred, truth = _codes(predictions, lut), _codes(ground_truth, lut)
    m = min(len(pred), len(truth))
    pred, truth = pred[:m], truth[:m]
    keep = (pred >= 0) & (truth >= 0)
    matrix = np.bincount(truth[keep] * n + pred[keep], minlength=n * n).reshape(n, n) if n else np.zeros((0, 0), int)
    return confusion_dict(matrix, labels)
