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
    min-max   max / min of a query's scores = Top-1 of q and of -q: a K = 1 search per modality with the query blocks
              [t], [-t] (resp. [i], [-i]) against that modality's column slice of the same fused database operand.
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
    """float64 column sums [D] and Gram matrix [D, D] of a (column slice of a) row-major fp32 matrix: one pass of the
    hand-written float64 contraction ``emr2a_gram_f64`` (no standardisation here: the rows are used as they are)."""
    n, d = int(unit_rows.shape[0]), int(unit_rows.shape[1])
    gram = torch.empty((d, d), dtype=torch.float64, device=eng.device)
    s = torch.empty((d,), dtype=torch.float64, device=eng.device)
    ws_bytes = int(eng.lib.emr2a_gram_f64_workspace_bytes(n, d))
    ws = torch.empty((ws_bytes // 8 + 2,), dtype=torch.float64, device=eng.device)
    ld = int(unit_rows.stride(0)) if n > 1 else d
    with torch.cuda.device(eng.device):
        native.check(eng.lib.emr2a_gram_f64(unit_rows.data_ptr(), ld, n, d, None, None, gram.data_ptr(), s.data_ptr(),
                                            ws.data_ptr(), ws.numel() * 8, eng._stream()))
    eng.launches += 2 if n else 0
    return s, gram


def _column_view(op: Operand, c0: int, c1: int) -> Operand:
    """Columns [c0, c1) of a prepared operand as an operand of its own (no copy): valid for the tensor-core arms
    when c0 and c1 are multiples of 64 (whole 128-byte k-chunks, nothing to zero-pad)."""
    cut = lambda t: None if t is None else t[:, c0:c1]               # noqa: E731
    return Operand(n=op.n, dim=c1 - c0, f32=cut(op.f32), hi=cut(op.hi), lo=cut(op.lo), stats=op.stats)


def _top1_scores(eng: Engine, q_rows: torch.Tensor, db: Operand, prec: str) -> Optional[torch.Tensor]:
    """Best score of every query row against ``db`` (float64), None if the rescore bound overflowed."""
    keys = eng.topk_search(_query_operand(eng, q_rows, prec), db, 1, prec)
    if prec == "rescore":
        _, overflow = eng.consume_status()
        if overflow:
            return None
    n_q = int(q_rows.shape[0])
    zeros_db = torch.zeros((db.n,), dtype=torch.int32, device=eng.device)
    zeros_q = torch.zeros((n_q,), dtype=torch.int32, device=eng.device)
    return eng.vote_metrics(keys, zeros_db, zeros_q, 1, k_list=[], per_query=False,
                            want_lists=True)["top_scores"][:, 0].double()


def _minmax_stats(eng: Engine, db: Operand, qu: torch.Tensor, d_t: int, prec: str):
    """(min, max - min + 1e-8) per query and modality: max / min of a query's scores = Top-1 of q and of -q.
    When the modality boundary falls on a k-chunk boundary each modality is searched against its own column
    slice of the fused operand (half the contraction length); otherwise the four zero-padded query blocks
    [t;0], [-t;0], [0;i], [0;-i] go against the fused rows."""
    n_q, dim = int(qu.shape[0]), int(qu.shape[1])
    out = []
    sliced = prec == "fp32" or (d_t % 64 == 0 and dim % 64 == 0)
    for c0, c1 in ((0, d_t), (d_t, dim)):
        if sliced:
            seg = qu[:, c0:c1]
            top = _top1_scores(eng, torch.cat([seg, -seg]).contiguous(), _column_view(db, c0, c1), prec)
        else:
            blocks = torch.zeros((2 * n_q, dim), dtype=torch.float32, device=eng.device)
            blocks[:n_q, c0:c1] = qu[:, c0:c1]
            blocks[n_q:, c0:c1] = -qu[:, c0:c1]
            top = _top1_scores(eng, blocks, db, prec)
        if top is None:
            return None
        lo = -top[n_q:]
        out.append((lo, top[:n_q] - lo + 1e-8))
    return out[0], out[1]


class LateFusionIndex:
    """A database prepared once for late-fusion searches (unit text rows | unit image rows + the planes of the search
    arm resident in HBM); the z-score moments (column sums + Gram matrix per modality, float64) are computed on first
    use and kept, so every later query batch only pays O(Q D^2) for its statistics."""

    def __init__(self, db_text, db_image, k: int = 10, precision: str = "auto", expected_queries: int = 4096,
                 engine: Optional[Engine] = None):
        self.eng = engine or get_engine()
        self.n_db, self.d_t, self.d_i = int(db_text.shape[0]), int(db_text.shape[1]), int(db_image.shape[1])
        if self.n_db == 0:
            raise ValueError("late fusion: empty database")
        if int(db_image.shape[0]) != self.n_db:
            raise ValueError("late fusion: text and image databases have different row counts")
        self.prec = self.eng.pick_precision(max(expected_queries, 1), self.n_db, self.d_t + self.d_i, k, precision)
        if self.prec == "bf16x1":
            raise ValueError("late fusion needs fp32-level scores (fp32, bf16x3 or rescore)")
        self.db = _db_operand(self.eng, db_text, db_image, self.prec)
        self._moments = None

    def moments(self):
        if self._moments is None:
            f32 = self.db.f32
            self._moments = (database_moments(self.eng, f32[:, :self.d_t]), database_moments(self.eng, f32[:, self.d_t:]))
        return self._moments

    def _zscore(self, moments, q_rows: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        s, gram = moments
        q64 = q_rows.double()
        mean = (q64 @ s) / self.n_db
        e2 = ((q64 @ gram) * q64).sum(dim=1) / self.n_db
        std = torch.sqrt(torch.clamp(e2 - mean * mean, min=0.0))
        return mean, std + 1e-8

    def search(self, q_text, q_image, text_weight: float, mode: int, k: int) -> Optional[torch.Tensor]:
        """Packed Top-k keys [Q, k]; None if the rescore bound overflowed (the caller rebuilds with bf16x3)."""
        eng, prec, d_t, d_i = self.eng, self.prec, self.d_t, self.d_i
        if prec == "rescore" and k > _RESCORE_MAX_K:
            raise ValueError(f"this index was built for the rescore arm (K <= {_RESCORE_MAX_K})")
        qu = eng.normalize_fuse(q_text, q_image, 1.0, 1.0, native.NF_SEGNORM, want_f32=True).f32      # [Q, Dt + Di]
        n_q = int(qu.shape[0])
        w64 = float(text_weight)
        if mode == native.SCORE_ZSCORE:
            m_t, m_i = self.moments()
            a_t, b_t = self._zscore(m_t, qu[:, :d_t])
            a_i, b_i = self._zscore(m_i, qu[:, d_t:])
        elif mode == native.SCORE_MINMAX:
            st = _minmax_stats(eng, self.db, qu, d_t, prec)
            if st is None:
                return None
            (a_t, b_t), (a_i, b_i) = st
        else:
            zero = torch.zeros((n_q,), dtype=torch.float64, device=eng.device)
            a_t, b_t, a_i, b_i = zero, zero + 1.0, zero, zero + 1.0
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
        keys = eng.topk_search(_query_operand(eng, scaled, prec), self.db, k, prec)
        if prec == "rescore":
            _, overflow = eng.consume_status()
            if overflow:
                return None
        if n_q:
            with torch.cuda.device(eng.device):
                native.check(eng.lib.emr2a_keys_add_offset(keys.data_ptr(), n_q, k, offset.data_ptr(), eng._stream()))
            eng.launches += 1
        return keys


def late_fusion_search(db_text, db_image, q_text, q_image, text_weight: float, mode: int, k: int,
                       precision: str = "auto", engine: Optional[Engine] = None) -> torch.Tensor:
    """Packed Top-k keys [Q, k] of  w * norm(cos_T) + (1 - w) * norm(cos_I)  per query, best first, scores = the
    fused (normalised) scores.  ``mode``: native.SCORE_ZSCORE / SCORE_MINMAX / SCORE_NONE.  One-shot form of
    ``LateFusionIndex`` (build + search)."""
    eng = engine or get_engine()
    index = LateFusionIndex(db_text, db_image, k, precision, max(int(q_text.shape[0]), 1), eng)
    keys = index.search(q_text, q_image, text_weight, mode, k)
    if keys is None:                              # rescore bound overflowed: the 3-pass arm is exact enough by itself
        del index
        index = LateFusionIndex(db_text, db_image, k, "bf16x3", max(int(q_text.shape[0]), 1), eng)
        keys = index.search(q_text, q_image, text_weight, mode, k)
    return keys
