"""Late fusion with per-query z-score / min-max score normalisation, without the [Q, N] score matrix
(retrieval/fusion.py:4-14, 31-42 and the late branch of RetrievalEvaluator.evaluate_retrieval,
retrieval/evaluator.py:150-157, which materialises every score of every query).

For one query the reference computes  fused = w * (ts - a_t) / b_t + (1 - w) * (is - a_i) / b_i  with
(a, b) = (mean, std + 1e-8) or (min, max - min + 1e-8) of that query's N cosine scores.  That is an affine
map of the two similarity vectors, so

    fused[d] = < [g_t * Tq ; g_i * Iq], [Td ; Id] > - c,    g_t = w / b_t,  g_i = (1 - w) / b_i,  c = g_t a_t + g_i a_i

and the Top-K is the ordinary fused search (K2) with per-row scaled query segments; the constant is applied to the
K winning scores afterwards.  The statistics never need the score matrix either:

    z-score   mean_q = <q, S> / N and E[s^2]_q = q^T G q / N from the database's column sums S and Gram matrix G
              (float64, one pass over the unit rows: emr2a_column_moments + a library DGEMM);
    min-max   max / min of a query's scores = Top-1 of q and of -q: one K = 1 search with the four query blocks
              [t;0], [-t;0], [0;i], [0;-i] against the same fused database operand.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import native
from . import preprocess as pp
from .engine import Engine, Operand, _ld, _RESCORE_MAX_K, get_engine

_CHUNK_ROWS = 1 << 18


def _db_operand(eng: Engine, db_text, db_image, prec: str) -> Operand:
    """Unit text rows | unit image rows, fp32 (statistics, re-scoring) + the planes the search arm needs."""
    return eng.normalize_fuse(db_text, db_image, 1.0, 1.0, native.NF_SEGNORM, want_f32=True,
                              want_planes=prec != "fp32", want_lo=prec == "bf16x3", want_stats=prec == "rescore")


def _query_operand(eng: Engine, rows: torch.Tensor, prec: str) -> Operand:
    return eng.prepare(rows, flags=0, precision=prec)


def database_moments(eng: Engine, unit_rows: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """float64 column sums [D] and Gram matrix [D, D] of a (column slice of a) row-major fp32 matrix."""
    n, d = int(unit_rows.shape[0]), int(unit_rows.shape[1])
    s, _ = pp.column_moments(eng, unit_rows)
    gram = torch.zeros((d, d), dtype=torch.float64, device=eng.device)
    for lo in range(0, n, _CHUNK_ROWS):
        z64 = unit_rows[lo:min(lo + _CHUNK_ROWS, n)].double()
        gram.addmm_(z64.t(), z64)
    return s, gram


def _zscore_stats(eng: Engine, db_rows: torch.Tensor, q_rows: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mean, std + 1e-8) of every query's N scores, float64 [Q]."""
    n = int(db_rows.shape[0])
    s, gram = database_moments(eng, db_rows)
    q64 = q_rows.double()
    mean = (q64 @ s) / n
    e2 = ((q64 @ gram) * q64).sum(dim=1) / n
    std = torch.sqrt(torch.clamp(e2 - mean * mean, min=0.0))
    return mean, std + 1e-8


def _minmax_stats(eng: Engine, db: Operand, qu: torch.Tensor, d_t: int, prec: str):
    """(min, max - min + 1e-8) per query and modality from one K = 1 search of [t;0], [-t;0], [0;i], [0;-i]."""
    n_q, dim = int(qu.shape[0]), int(qu.shape[1])
    blocks = torch.zeros((4 * n_q, dim), dtype=torch.float32, device=eng.device)
    blocks[0 * n_q:1 * n_q, :d_t] = qu[:, :d_t]
    blocks[1 * n_q:2 * n_q, :d_t] = -qu[:, :d_t]
    blocks[2 * n_q:3 * n_q, d_t:] = qu[:, d_t:]
    blocks[3 * n_q:4 * n_q, d_t:] = -qu[:, d_t:]
    keys = eng.topk_search(_query_operand(eng, blocks, prec), db, 1, prec)
    if prec == "rescore":
        _, overflow = eng.consume_status()
        if overflow:
            return None
    top = eng.vote_metrics(keys, torch.zeros((db.n,), dtype=torch.int32, device=eng.device),
                           torch.zeros((4 * n_q,), dtype=torch.int32, device=eng.device), 1, k_list=[],
                           per_query=False, want_lists=True)["top_scores"][:, 0].double()
    t_max, t_min = top[0 * n_q:1 * n_q], -top[1 * n_q:2 * n_q]
    i_max, i_min = top[2 * n_q:3 * n_q], -top[3 * n_q:4 * n_q]
    return (t_min, t_max - t_min + 1e-8), (i_min, i_max - i_min + 1e-8)


def late_fusion_search(db_text, db_image, q_text, q_image, text_weight: float, mode: int, k: int,
                       precision: str = "auto", engine: Optional[Engine] = None) -> torch.Tensor:
    """Packed Top-k keys [Q, k] of  w * norm(cos_T) + (1 - w) * norm(cos_I)  per query, best first, scores = the
    fused (normalised) scores.  ``mode``: native.SCORE_ZSCORE / SCORE_MINMAX / SCORE_NONE."""
    eng = engine or get_engine()
    n_db, d_t, d_i = int(db_text.shape[0]), int(db_text.shape[1]), int(db_image.shape[1])
    n_q = int(q_text.shape[0])
    if n_db == 0:
        raise ValueError("late_fusion_search: empty database")
    prec = eng.pick_precision(max(n_q, 1), n_db, d_t + d_i, k, precision)
    if prec == "bf16x1":
        raise ValueError("late_fusion_search needs fp32-level scores (fp32, bf16x3 or rescore)")
    db = _db_operand(eng, db_text, db_image, prec)
    qu = eng.normalize_fuse(q_text, q_image, 1.0, 1.0, native.NF_SEGNORM, want_f32=True).f32      # [Q, Dt + Di]
    w64 = float(text_weight)
    if mode == native.SCORE_ZSCORE:
        a_t, b_t = _zscore_stats(eng, db.f32[:, :d_t], qu[:, :d_t])
        a_i, b_i = _zscore_stats(eng, db.f32[:, d_t:], qu[:, d_t:])
    elif mode == native.SCORE_MINMAX:
        st = _minmax_stats(eng, db, qu, d_t, prec)
        if st is None:                                            # rescore bound overflowed: exact-enough 3-pass arm
            return late_fusion_search(db_text, db_image, q_text, q_image, text_weight, mode, k, "bf16x3", eng)
        (a_t, b_t), (a_i, b_i) = st
    else:
        zero, one = torch.zeros((n_q,), dtype=torch.float64, device=eng.device), torch.ones((n_q,), dtype=torch.float64, device=eng.device)
        a_t, b_t, a_i, b_i = zero, one, zero, one
    # the reference applies mean / std (min / range) as fp32 scalars: (scores - f32(a)) / f32(b), then w * (.)
    a_t, b_t = a_t.float().double(), b_t.float().double()
    a_i, b_i = a_i.float().double(), b_i.float().double()
    g_t = float(np.float32(w64)) / b_t
    g_i = float(np.float32(1.0 - w64)) / b_i
    offset = (-(g_t * a_t + g_i * a_i)).float().contiguous()
    g_t32, g_i32 = g_t.float().contiguous(), g_i.float().contiguous()
    scaled = qu.clone()
    if n_q:
        with torch.cuda.device(eng.device):
            native.check(eng.lib.emr2a_scale_segments(scaled.data_ptr(), n_q, d_t, d_i, _ld(scaled), g_t32.data_ptr(),
                                                      g_i32.data_ptr(), eng._stream()))
        eng.launches += 1
    keys = eng.topk_search(_query_operand(eng, scaled, prec), db, k, prec)
    if prec == "rescore":
        _, overflow = eng.consume_status()
        if overflow:
            return late_fusion_search(db_text, db_image, q_text, q_image, text_weight, mode, k, "bf16x3", eng)
    if n_q:
        with torch.cuda.device(eng.device):
            native.check(eng.lib.emr2a_keys_add_offset(keys.data_ptr(), n_q, k, offset.data_ptr(), eng._stream()))
        eng.launches += 1
    return keys
