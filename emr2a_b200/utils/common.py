"""``utils.common`` on the B200 (reference: utils/common.py:4-22): single-vector
normalise with a zero guard (no epsilon) and text-first weighted concat -- K1 with n = 1."""
import numpy as np

from .. import native
from ..engine import get_engine


def l2_normalize(vec: np.ndarray) -> np.ndarray:
    vec = np.asarray(vec)
    out = get_engine().normalize_fuse(vec.reshape(1, -1), flags=native.NF_ZERO_GUARD).f32[0].cpu().numpy()
    return out.astype(vec.dtype, copy=False) if vec.dtype == np.float64 else out


def concat_embeddings(text_emb: np.ndarray, image_emb: np.ndarray, text_weight: float = 1.0,
                      image_weight: float = 1.0) -> np.ndarray:
    text_emb = np.asarray(text_emb)
    image_emb = np.asarray(image_emb)
    op = get_engine().normalize_fuse(text_emb.reshape(1, -1), image_emb.reshape(1, -1),
                                     w0=np.float32(float(text_weight)), w1=np.float32(float(image_weight)),
                                     flags=native.NF_ZERO_GUARD)
    return op.f32[0].cpu().numpy()
