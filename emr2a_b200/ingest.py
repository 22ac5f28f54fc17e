"""Embedding ingest for the array fast path.

The reference's callers keep embeddings as python dicts and loop over patients:
  * ``pipelines/step3_retrieval/evaluate_retrieval.py:30-33, 66-67``: ``.npz`` with one key per patient,
    value ``(n_slices, D)``, mean-pooled with ``embeddings[pid].mean(axis=0)`` per patient;
  * ``analysis/run_cv_experiments.py:111-128``: ``.npz`` with ``patient_ids`` / ``image_matrix`` /
    ``text_matrix`` (re-decompressed once per patient -- O(N^2)), ``aggregate_embeddings`` (:316-333).
Those scripts run unchanged on the drop-in evaluators; this module is the dict-free equivalent for large
runs: one pass over the file, one H2D copy from pinned memory, slice mean-pool on the GPU
(``emr2a_segment_mean``), arrays out.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import get_engine


def mean_pool_patients(per_patient: Sequence[np.ndarray]) -> torch.Tensor:
    """[(n_slices_p, D) or (D,)] per patient -> device tensor [n_patients, D] of slice means
    (``arr.mean(axis=0)`` arithmetic: slices added in order in fp32, divided by the count)."""
    eng = get_engine()
    mats = [np.atleast_2d(np.asarray(a, dtype=np.float32)) for a in per_patient]
    if not mats:
        return torch.empty((0, 0), dtype=torch.float32, device=eng.device)
    counts = np.fromiter((m.shape[0] for m in mats), dtype=np.int64, count=len(mats))
    offsets = np.concatenate([[0], np.cumsum(counts)])
    flat = torch.from_numpy(np.concatenate(mats, axis=0)).pin_memory()
    return eng.segment_mean(flat.to(eng.device, non_blocking=True), offsets)


def load_patient_npz(path, patient_ids: Optional[Sequence[str]] = None) -> Tuple[List[str], torch.Tensor]:
    """Step-2 layout (one key per patient, value (n_slices, D)): returns (ids, pooled [n, D] on the device).
    Each array is read from the archive exactly once."""
    with np.load(path) as data:
        ids = list(patient_ids) if patient_ids is not None else list(data.files)
        mats = [data[pid] for pid in ids]
    return ids, mean_pool_patients(mats)


def load_matrix_npz(path) -> Dict[str, object]:
    """CV-runner layout (``patient_ids`` + ``image_matrix`` and/or ``text_matrix``): every member is
    decompressed once (the reference re-reads the whole matrix per patient)."""
    out: Dict[str, object] = {}
    with np.load(path, allow_pickle=True) as data:
        out["patient_ids"] = [str(p) for p in data["patient_ids"]]
        for key, name in (("image_matrix", "image"), ("text_matrix", "text")):
            if key in data.files:
                out[name] = np.ascontiguousarray(data[key], dtype=np.float32)
    return out


def embeddings_to_arrays(patient_ids: Sequence[str], embeddings: Dict[str, Dict[str, np.ndarray]]
                         ) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
    """The ``run_cv`` embeddings dict -> (image [n, Di], text [n, Dt]) stacked ONCE in patient order
    (the reference re-stacks per fold, utils/cv_evaluator.py:366-371)."""
    def stack(modality):
        if not patient_ids or modality not in embeddings[patient_ids[0]]:
            return None
        return np.stack([embeddings[p][modality] for p in patient_ids])
    return stack("image"), stack("text")
