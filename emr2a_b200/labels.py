"""String labels / ids <-> integer codes, and vectorised materialisation of the
reference's list-of-lists result structures."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def encode(*label_lists: Sequence) -> Tuple[List[str], List[np.ndarray]]:
    """classes = sorted(set(all labels)) (utils/cv_evaluator.py:312); returns the class
    list and one int32 code array per input list."""
    classes = sorted(set(lab for labs in label_lists for lab in labs))
    lut = {c: i for i, c in enumerate(classes)}
    return classes, [np.fromiter((lut[l] for l in labs), dtype=np.int32, count=len(labs)) for labs in label_lists]


def gather_lists(table: Sequence, idx: np.ndarray, valid: np.ndarray) -> List[list]:
    """[[table[j] for j in row[:valid_r]] for each row] without a python inner loop
    when all rows are full."""
    arr = np.asarray(table, dtype=object)
    if idx.size == 0:
        return [[] for _ in range(idx.shape[0])]
    safe = np.where(idx >= 0, idx, 0)
    picked = arr[safe]
    if valid.min(initial=idx.shape[1]) == idx.shape[1]:
        return picked.tolist()
    return [picked[r, :valid[r]].tolist() for r in range(idx.shape[0])]


def score_lists(scores: np.ndarray, valid: np.ndarray) -> List[List[float]]:
    """fp32 scores widened to python floats, as ``float(similarities[i])`` does
    (utils/cv_evaluator.py:125)."""
    wide = scores.astype(np.float64)
    if wide.size and valid.min(initial=scores.shape[1]) == scores.shape[1]:
        return wide.tolist()
    return [wide[r, :valid[r]].tolist() for r in range(scores.shape[0])]
