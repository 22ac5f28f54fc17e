"""Per-fold preprocessing on the B200: StandardScaler -> PCA fitted on the train fold, applied to both
sides (``CVRetrievalEvaluator.process_embeddings`` utils/cv_evaluator.py:73-93; the ``use_pca`` branch of
``RetrievalEvaluator.process_embeddings`` retrieval/evaluator.py:44-73).  SURVEY §8f-3.

What runs where
    column mean / variance ........ ``emr2a_column_moments`` (hand-written, HBM-bound, float64 accumulators)
    (x - mean) / scale ............ ``emr2a_standardize``    (hand-written, HBM-bound, sklearn's fp32 arithmetic)
    Z^T Z, column sums of Z ....... ``emr2a_gram_f64`` (hand-written float64 FMA contraction, standardisation fused
                                    into its operand load: the raw train rows are read, Z is never written)
    eigh .......................... library call (cuSOLVER syevd through torch): a D x D symmetric eigenproblem, fit only
    Z W^T - bias .................. ``emr2a_project`` (hand-written fp32 FMA GEMM, fixed summation order, standardisation
                                    fused into its operand load, bias into its epilogue)
    scaler-only transform ......... K1 with ``NF_STANDARDIZE`` (standardise + row-normalise in one pass)
    row normalisation ............. K1 (``emr2a_normalize_fuse``)

Contract.  The scaler reproduces sklearn's ``StandardScaler`` (float64 statistics, fp32 transform).  The PCA is the
EXACT principal-component basis of the standardised train rows: float64 covariance, float64 symmetric
eigen-decomposition, components ordered by decreasing eigenvalue, signs fixed by sklearn's
``svd_flip(u_based_decision=False)`` rule (largest-|.| entry of every component is positive).  That is what
``PCA(svd_solver="full")`` / ``"covariance_eigh"`` compute up to their own fp32 rounding.  (The sign rule is
ill-conditioned when the two largest entries of an axis tie -- every 2-feature fold has the axes (1, 1)/sqrt(2),
(1, -1)/sqrt(2) -- and LAPACK / cuSOLVER may break such a tie differently; the sign of an axis flips train and test
rows alike and leaves every cosine score unchanged.)  sklearn's ``"auto"``
picks the *randomized* solver for mid-sized folds (e.g. 1600 x 512 -> 128) and the reference does not seed it, so
the reference's own output there differs from run to run; the exact basis is deterministic.

There is no CPU fallback: everything here needs the CUDA engine.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import native
from .engine import Engine, _ld, get_engine

_CHUNK_ROWS = 1 << 18          # rows standardised / multiplied at a time (bounds the fp64 staging copy)
_F64_EPS = float(np.finfo(np.float64).eps)


@dataclass
class FoldTransform:
    """Fitted StandardScaler (+ PCA) of one train fold; all tensors live on the engine's device."""
    n_train: int
    dim: int
    mean: torch.Tensor                 # float64 [D]   StandardScaler.mean_
    var: torch.Tensor                  # float64 [D]   StandardScaler.var_
    scale: torch.Tensor                # float64 [D]   StandardScaler.scale_ (constant features -> 1)
    mean_f32: torch.Tensor             # what transform() subtracts / divides by
    scale_f32: torch.Tensor
    components: Optional[torch.Tensor] = None          # float32 [P, D]  PCA.components_
    pca_mean: Optional[torch.Tensor] = None            # float32 [D]     PCA.mean_ (mean of the standardised rows)
    bias: Optional[torch.Tensor] = None                # float32 [P]     pca_mean @ components^T
    explained_variance: Optional[torch.Tensor] = None  # float64 [P]

    @property
    def n_components(self) -> int:
        return 0 if self.components is None else int(self.components.shape[0])


class _NoTF32:
    """fp32 library GEMMs must be true fp32 here, whatever the process-wide torch setting is."""
    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


def column_moments(eng: Engine, x: torch.Tensor, shift: Optional[torch.Tensor] = None
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """float64 (sum, sum of squares) per column of ``x - shift``."""
    n, d = int(x.shape[0]), int(x.shape[1])
    s = torch.empty((d,), dtype=torch.float64, device=eng.device)
    ss = torch.empty((d,), dtype=torch.float64, device=eng.device)
    ws_bytes = int(eng.lib.emr2a_column_moments_workspace_bytes(n, d))
    ws = torch.empty((ws_bytes // 8 + 2,), dtype=torch.float64, device=eng.device)
    with torch.cuda.device(eng.device):
        native.check(eng.lib.emr2a_column_moments(x.data_ptr(), _ld(x), n, d, native.ptr(shift), s.data_ptr(),
                                                  ss.data_ptr(), ws.data_ptr(), ws.numel() * 8, eng._stream()))
    eng.launches += 2 if n else 0
    return s, ss


def standardize(eng: Engine, x: torch.Tensor, mean_f32: torch.Tensor, scale_f32: torch.Tensor,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, d = int(x.shape[0]), int(x.shape[1])
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=eng.device)
    if n:
        with torch.cuda.device(eng.device):
            native.check(eng.lib.emr2a_standardize(x.data_ptr(), _ld(x), n, d, mean_f32.data_ptr(), scale_f32.data_ptr(),
                                                   out.data_ptr(), _ld(out), eng._stream()))
        eng.launches += 1
    return out


def fit_scaler(eng: Engine, x: torch.Tensor) -> FoldTransform:
    """StandardScaler.fit on device rows (sklearn: float64 mean / population variance; a feature whose variance
    is within rounding of zero gets scale 1 -- ``_is_constant_feature`` / ``_handle_zeros_in_scale``)."""
    n, d = int(x.shape[0]), int(x.shape[1])
    if n == 0:
        raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by StandardScaler.")
    shift = x[0].contiguous()                        # any row: keeps sum((x - shift)^2) free of cancellation
    s, ss = column_moments(eng, x, shift)
    dm = s / n
    mean = shift.double() + dm
    var = torch.clamp(ss / n - dm * dm, min=0.0)
    upper = n * _F64_EPS * var + (n * mean * _F64_EPS) ** 2
    scale = torch.sqrt(var)
    scale = torch.where((var <= upper) | (scale == 0.0), torch.ones_like(scale), scale)
    return FoldTransform(n_train=n, dim=d, mean=mean, var=var, scale=scale,
                         mean_f32=mean.float().contiguous(), scale_f32=scale.float().contiguous())


def fit(x_train, pca_dim: Optional[int], engine: Optional[Engine] = None) -> FoldTransform:
    """Fit scaler and (if ``min(pca_dim, n - 1, D) > 0``) the exact PCA on the train rows."""
    eng = engine or get_engine()
    x, _ = eng._embedding(x_train)
    if x.dtype != torch.float32:
        x = x.float()
    tf = fit_scaler(eng, x)
    n, d = tf.n_train, tf.dim
    n_comp = min(int(pca_dim), n - 1, d) if pca_dim is not None else 0
    if n_comp <= 0:
        return tf
    # covariance of the standardised rows, float64: C = (Z^T Z - n m m^T) / (n - 1).  Z^T Z and the column sums come
    # from ONE pass over the raw rows (emr2a_gram_f64: standardisation fused into the operand load, float64 FMA
    # contraction, upper triangle only); the standardised matrix is never written.
    gram = torch.empty((d, d), dtype=torch.float64, device=eng.device)
    zsum = torch.empty((d,), dtype=torch.float64, device=eng.device)
    ws_bytes = int(eng.lib.emr2a_gram_f64_workspace_bytes(n, d))
    ws = torch.empty((ws_bytes // 8 + 2,), dtype=torch.float64, device=eng.device)
    with torch.cuda.device(eng.device):
        native.check(eng.lib.emr2a_gram_f64(x.data_ptr(), _ld(x), n, d, tf.mean_f32.data_ptr(), tf.scale_f32.data_ptr(),
                                            gram.data_ptr(), zsum.data_ptr(), ws.data_ptr(), ws.numel() * 8, eng._stream()))
    eng.launches += 2
    m = zsum / n
    cov = (gram - n * torch.outer(m, m)) / (n - 1)
    cov = 0.5 * (cov + cov.t())
    evals, evecs = torch.linalg.eigh(cov)                       # ascending
    evals = torch.flip(evals, dims=(0,))[:n_comp].clamp_(min=0.0)
    comps = torch.flip(evecs, dims=(1,)).t()[:n_comp].contiguous()        # [P, D], decreasing eigenvalue
    pivot = comps.abs().argmax(dim=1)
    signs = torch.sign(comps[torch.arange(n_comp, device=eng.device), pivot])
    comps = comps * signs[:, None]
    tf.components = comps.float().contiguous()
    tf.pca_mean = m.float().contiguous()
    tf.bias = eng.scores(tf.pca_mean[None, :].contiguous(), tf.components)[0].contiguous()      # mean_ @ components_.T
    tf.explained_variance = evals
    return tf


def transform(tf: FoldTransform, x, engine: Optional[Engine] = None, normalize: bool = True) -> torch.Tensor:
    """Scaler (+ PCA projection ``Z W^T - mean W^T``, PCA._transform) (+ row L2 normalisation, K1) of device or
    host rows; returns a device tensor [n, P or D]."""
    eng = engine or get_engine()
    x, _ = eng._embedding(x)
    if x.dtype != torch.float32:
        x = x.float()
    n, d = int(x.shape[0]), int(x.shape[1])
    if d != tf.dim:
        raise ValueError(f"X has {d} features, but the transform was fitted with {tf.dim} features")
    if tf.components is None:
        if normalize:
            # scaler-only transform: standardise + row-normalise in ONE pass over the raw rows (K1, NF_STANDARDIZE)
            col_std = torch.stack([tf.mean_f32, tf.scale_f32, 1.0 / tf.scale_f32]).contiguous()
            try:
                return eng.normalize_fuse(x, flags=native.NF_ROWNORM | native.NF_STANDARDIZE, col_std=col_std).f32
            except native.Emr2aError as exc:      # shapes the fused variant does not take (odd / very wide rows)
                if exc.code != native.ERR_UNSUPPORTED:
                    raise
        y = standardize(eng, x, tf.mean_f32, tf.scale_f32)
    else:
        # projection with the standardisation fused into the operand load and the bias into the epilogue
        # (emr2a_project: one kernel, the raw rows are read once, nothing but the [n, P] result is written)
        p = tf.n_components
        y = torch.empty((n, p), dtype=torch.float32, device=eng.device)
        w = tf.components                                   # [P, D]: y[r, j] = <z_r, w_j> - bias_j
        if n:
            with torch.cuda.device(eng.device):
                native.check(eng.lib.emr2a_project(x.data_ptr(), _ld(x), n, d, tf.mean_f32.data_ptr(), tf.scale_f32.data_ptr(),
                                                   w.data_ptr(), _ld(w), p, tf.bias.data_ptr(), y.data_ptr(), p, eng._stream()))
            eng.launches += 1
    if not normalize:
        return y
    return eng.normalize_fuse(y, flags=native.NF_ROWNORM).f32


def sklearn_solver(n_samples: int, n_features: int, n_components: int) -> str:
    """The solver ``PCA(n_components)`` with ``svd_solver="auto"`` picks (sklearn 1.5+ ``PCA._fit``)."""
    if n_features <= 1000 and n_samples >= 10 * n_features:
        return "covariance_eigh"
    if max(n_samples, n_features) <= 500:
        return "full"
    if 1 <= n_components < 0.8 * min(n_samples, n_features):
        return "randomized"
    return "full"
