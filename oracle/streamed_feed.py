"""TEST INFRASTRUCTURE ONLY (see emr2a_oracle.py): feeds ``StreamedReferenceSample`` with a database that lives on
the GPU (or is generated there chunk by chunk), so that the oracle can check a sample of queries against ALL rows of
the 1M-10M row configurations without the host ever holding the database.  Only ``tests/`` and the ``cpu_baseline``
leg of ``bench.py`` import this."""
from __future__ import annotations

import concurrent.futures as cf
import os
from typing import Callable, Optional, Tuple

import numpy as np


def feed(sample, n_rows: int, fetch: Callable[[int, int], Tuple], chunk_rows: int = 65536,
         timed_rows: int = 262144, workers: Optional[int] = None) -> None:
    """``fetch(row0, rows) -> (img, txt | None, fold | None)`` as host numpy arrays.  The first ``timed_rows`` rows are
    fed the reference's way, serially and timed (the bounded CPU-baseline sample); the remaining chunks are scored
    with one sgemm each on a thread pool (parity only)."""
    workers = workers or max(1, min(8, (os.cpu_count() or 2) - 1))
    row0 = 0
    while row0 < min(timed_rows, n_rows):
        rows = min(chunk_rows, n_rows - row0)
        img, txt, fold = fetch(row0, rows)
        sample.add_chunk(row0, img, txt, fold, timed=True)
        row0 += rows
    pending = []
    with cf.ThreadPoolExecutor(max_workers=workers) as pool:
        while row0 < n_rows:
            rows = min(chunk_rows, n_rows - row0)
            img, txt, fold = fetch(row0, rows)
            pending.append(pool.submit(sample.add_chunk, row0, img, txt, fold, False))
            row0 += rows
            while len(pending) > workers + 1:
                pending.pop(0).result()
        for fut in pending:
            fut.result()


def device_fetcher(img_dev, txt_dev=None, fold_dev=None):
    """Rows of device-resident matrices (any float dtype; bf16 is widened to fp32 exactly) as host arrays."""
    import torch

    def fetch(row0: int, rows: int):
        def host(t):
            return None if t is None else t[row0:row0 + rows].to(torch.float32).cpu().numpy()
        fold = None if fold_dev is None else fold_dev[row0:row0 + rows].cpu().numpy()
        return host(img_dev), host(txt_dev), fold
    return fetch


def generated_fetcher(gen_img, gen_txt=None, fold_of=None):
    """Rows regenerated on the device per chunk: ``gen_*(row0, rows) -> device tensor``; ``fold_of(row0, rows) ->
    numpy uint8`` -- for shards that are spread over several GPUs (the 8-GPU C5 run)."""
    import torch

    def fetch(row0: int, rows: int):
        img = gen_img(row0, rows).to(torch.float32).cpu().numpy()
        txt = None if gen_txt is None else gen_txt(row0, rows).to(torch.float32).cpu().numpy()
        return img, txt, (None if fold_of is None else fold_of(row0, rows))
    return fetch
