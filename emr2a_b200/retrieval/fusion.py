"""``retrieval.fusion`` on the B200: same signatures as the reference
(retrieval/fusion.py:4-42)."""
import numpy as np

from .. import native
from ..engine import get_engine

_MODES = {"none": native.SCORE_NONE, "zscore": native.SCORE_ZSCORE, "minmax": native.SCORE_MINMAX}


def late_fusion(text_scores: np.ndarray, image_scores: np.ndarray, text_weight: float = 0.4,
                score_mode: str = "none") -> np.ndarray:
    """``w * norm(text) + (1 - w) * norm(image)`` per query row (retrieval/fusion.py:4-14).
    Accepts one score vector (N,) or a batch (Q, N)."""
    eng = get_engine()
    ts = np.asarray(text_scores, dtype=np.float32)
    im = np.asarray(image_scores, dtype=np.float32)
    if ts.shape != im.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {ts.shape} {im.shape}")
    mode = _MODES.get(score_mode, native.SCORE_NONE)       # unknown modes pass scores through, as the reference does
    return eng.late_fuse_scores(ts, im, text_weight, mode).cpu().numpy()


def early_fusion(text_embeddings: np.ndarray, image_embeddings: np.ndarray, text_weight: float = 1.0,
                 image_weight: float = 1.0) -> np.ndarray:
    """Text-first weighted concatenation + row L2 normalisation in one K1 pass
    (retrieval/fusion.py:17-28)."""
    eng = get_engine()
    op = eng.normalize_fuse(np.asarray(text_embeddings), np.asarray(image_embeddings),
                            w0=np.float32(text_weight), w1=np.float32(image_weight), flags=native.NF_ROWNORM)
    return op.f32.cpu().numpy()


def normalize_scores(scores: np.ndarray, mode: str = "none") -> np.ndarray:
    """none / zscore / minmax over the scores of one query (retrieval/fusion.py:31-42)."""
    if mode not in ("zscore", "minmax"):
        return scores
    eng = get_engine()
    s = np.asarray(scores, dtype=np.float32)
    # 1 * norm(s) + 0 * norm(s) == norm(s) exactly
    return eng.late_fuse_scores(s, s, 1.0, _MODES[mode]).cpu().numpy()
