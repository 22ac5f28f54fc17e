"""Seeded synthetic embeddings shaped like the BASELINE.json configs.

Class-structured Gaussians ``x = s * mu[label] + eps``: with
``s = 0.08`` the votes are non-trivial (top-1 ~ 0.5, vote ~ 0.6 at C1) instead
of saturating at 1.0.  Everything is generated with ``numpy.random.default_rng``
on the host; the multi-GPU bench generates per-shard blocks with the same
formula on the device (see ``device_block``) so a 41 GB database never sits in
host memory.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

SEP = 0.08


def class_gaussians(n: int, dim: int, n_classes: int, seed: int, sep: float = SEP,
                    labels: np.ndarray | None = None) -> Tuple[np.ndarray, np.ndarray]:
    """Return (embeddings float32 [n, dim], labels int32 [n])."""
    rng = np.random.default_rng(seed)
    if labels is None:
        labels = rng.integers(0, n_classes, size=n).astype(np.int32)
    centres = rng.standard_normal((n_classes, dim)).astype(np.float32)
    x = rng.standard_normal((n, dim), dtype=np.float32)
    x += np.float32(sep) * centres[labels]
    return x, labels


def two_modal(n: int, d_img: int, d_txt: int, n_classes: int, seed: int,
              sep: float = SEP) -> Dict[str, np.ndarray]:
    """Image + text embeddings sharing one label vector."""
    img, labels = class_gaussians(n, d_img, n_classes, seed, sep)
    txt, _ = class_gaussians(n, d_txt, n_classes, seed + 1, sep, labels=labels)
    return {"image": img, "text": txt, "labels": labels}


def patient_ids(n: int) -> list:
    return [f"p{i:07d}" for i in range(n)]


def label_names(codes: np.ndarray, n_classes: int) -> list:
    names = [f"class_{c}" for c in range(n_classes)]
    return [names[int(c)] for c in codes]


def device_block(row0: int, rows: int, dim: int, n_classes: int, seed: int, device,
                 sep: float = SEP, label_seed: int | None = None, dtype=None):
    """Generate rows [row0, row0+rows) of a (virtually unbounded) class-structured
    database directly on ``device``.  Deterministic per (seed, 65536-row chunk): the
    generators are re-seeded per chunk so any shard layout reproduces the same rows.
    Labels come from ``label_seed`` (default ``seed``) so several modalities can share
    one label vector.  Returns (float32 [rows, dim], int32 labels [rows]); ``dtype`` (e.g. torch.bfloat16)
    rounds the fp32 values chunk by chunk into a pre-allocated result (no fp32 copy of the whole block)."""
    import torch

    chunk = 65536
    label_seed = seed if label_seed is None else label_seed
    out = torch.empty((rows, dim), dtype=dtype or torch.float32, device=device)
    ls = []
    g = torch.Generator(device=device)
    gl = torch.Generator(device=device)
    gc = torch.Generator(device=device)
    gc.manual_seed(seed * 1000003 + 17)
    centres = torch.randn((n_classes, dim), generator=gc, device=device, dtype=torch.float32)
    r = row0
    end = row0 + rows
    while r < end:
        c0 = (r // chunk) * chunk
        gl.manual_seed(label_seed * 1000003 + 7919 + c0 // chunk)
        g.manual_seed(seed * 1000003 + 101 + c0 // chunk)
        lab = torch.randint(0, n_classes, (chunk,), generator=gl, device=device, dtype=torch.int32)
        x = torch.randn((chunk, dim), generator=g, device=device, dtype=torch.float32)
        x += sep * centres[lab.long()]
        lo, hi = r - c0, min(end, c0 + chunk) - c0
        out[r - row0:r - row0 + (hi - lo)] = x[lo:hi]
        ls.append(lab[lo:hi])
        r = c0 + hi
    return out, (torch.cat(ls) if ls else torch.empty((0,), dtype=torch.int32, device=device))


def device_labels(row0: int, rows: int, n_classes: int, label_seed: int, device):
    """Labels of rows [row0, row0+rows) as ``device_block`` assigns them."""
    import torch

    chunk = 65536
    gl = torch.Generator(device=device)
    out = []
    r, end = row0, row0 + rows
    while r < end:
        c0 = (r // chunk) * chunk
        gl.manual_seed(label_seed * 1000003 + 7919 + c0 // chunk)
        lab = torch.randint(0, n_classes, (chunk,), generator=gl, device=device, dtype=torch.int32)
        lo, hi = r - c0, min(end, c0 + chunk) - c0
        out.append(lab[lo:hi])
        r = c0 + hi
    return torch.cat(out)
