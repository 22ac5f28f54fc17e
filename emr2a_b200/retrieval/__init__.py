"""B200 implementation of the reference's ``retrieval`` package (same public names,
retrieval/__init__.py:1-11)."""
from .similarity import compute_cosine_similarity, compute_euclidean_similarity
from .fusion import late_fusion, early_fusion
from .evaluator import RetrievalEvaluator

__all__ = [
    "compute_cosine_similarity",
    "compute_euclidean_similarity",
    "late_fusion",
    "early_fusion",
    "RetrievalEvaluator",
]
