"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: shard ranges, the all-gather of packed
Top-K keys, and the property that merging per-shard Top-K lists by key equals the global Top-K
(so results cannot depend on the GPU count)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def _pack(scores, idx):
    b = scores.astype(np.float32).view(np.uint32).astype(np.uint64)
    o = np.where(b & np.uint64(0x80000000), (~b) & np.uint64(0xFFFFFFFF), b ^ np.uint64(0x80000000))
    return ((o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - idx.astype(np.uint64))).view(np.int64)


def _local_topk(scores, lo, k):
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return _pack(np.take_along_axis(scores, order, axis=1), order + lo)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emr2a_b200.dist import gather_keys, shard_range
    rng = np.random.default_rng(0)                      # same data on every rank
    n, n_q, k = 1000, 17, 5
    scores = rng.standard_normal((n_q, n)).astype(np.float32)
    scores[:, 700] = scores[:, 100]                     # ties across the shard boundary
    lo, hi = shard_range(n, rank, world, align=256)
    local = torch.from_numpy(_local_topk(scores[:, lo:hi], lo, k))
    allk = gather_keys(local).numpy()                   # [world, Q, K]
    flat = np.transpose(allk, (1, 0, 2)).reshape(n_q, -1).view(np.uint64)
    merged = np.sort(flat, axis=1)[:, ::-1][:, :k]          # what the K3 kernel computes on the GPU
    want = _local_topk(scores, 0, k).view(np.uint64)
    q.put((rank, bool(np.array_equal(merged, want)), (lo, hi)))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from emr2a_b200.dist import shard_range
    for n in (0, 1, 255, 256, 1000, 1_000_000, 10_000_001):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all(lo % 256 == 0 or lo == n for lo, _ in spans)


def test_gather_and_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [o[1] for o in out] == [True, True]
    assert out[0][2] == (0, 512) and out[1][2] == (512, 1000)
