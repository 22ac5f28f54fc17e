"""``retrieval.similarity`` on the B200: same signatures as the reference
(retrieval/similarity.py:4-15), computed by libemr2a.so kernels."""
import numpy as np

from .. import native
from ..engine import get_engine


def _out_dtype(*arrays):
    dt = np.result_type(*[np.asarray(a).dtype for a in arrays])
    return dt if dt in (np.float32, np.float64) else np.float32


def compute_cosine_similarity(query: np.ndarray, database: np.ndarray) -> np.ndarray:
    """Cosine of ``query`` (D,) against every row of ``database`` (N, D).

    Both sides are divided by (norm + 1e-8) by the K1 kernel, the dot products
    come from the fp32 score kernel (reference: retrieval/similarity.py:4-7).
    """
    eng = get_engine()
    query = np.asarray(query)
    database = np.asarray(database)
    if database.ndim != 2 or query.ndim != 1 or database.shape[1] != query.shape[0]:
        raise ValueError(f"shapes {query.shape} and {database.shape} not aligned")
    db = eng.normalize_fuse(database, flags=native.NF_ROWNORM)
    q = eng.normalize_fuse(query[None, :], flags=native.NF_ROWNORM)
    out = eng.scores(q.f32, db.f32)[0]
    return out.cpu().numpy().astype(_out_dtype(query, database), copy=False)


def compute_euclidean_similarity(query: np.ndarray, database: np.ndarray) -> np.ndarray:
    """``1 - dist / max(dist)`` (reference: retrieval/similarity.py:10-15)."""
    eng = get_engine()
    query = np.asarray(query)
    database = np.asarray(database)
    if database.ndim != 2 or query.ndim != 1 or database.shape[1] != query.shape[0]:
        raise ValueError(f"shapes {query.shape} and {database.shape} not aligned")
    out = eng.euclid_scores(query.astype(np.float32, copy=False), database.astype(np.float32, copy=False))
    return out.cpu().numpy().astype(_out_dtype(query, database), copy=False)
