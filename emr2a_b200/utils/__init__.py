"""B200 implementation of the reference's ``utils`` package surface that lies on the
retrieval hot path (utils/__init__.py:1-16): ``common``, ``metrics``, ``cv_evaluator``."""
from .common import l2_normalize, concat_embeddings
from .metrics import (
    compute_accuracy,
    compute_top_k_accuracy,
    compute_precision_recall_f1,
    compute_confusion_matrix,
)

__all__ = [
    "l2_normalize",
    "concat_embeddings",
    "compute_accuracy",
    "compute_top_k_accuracy",
    "compute_precision_recall_f1",
    "compute_confusion_matrix",
]
