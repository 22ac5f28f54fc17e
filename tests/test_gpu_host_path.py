"""Engine.search_and_vote_host -- the end-to-end call bench.py times as `e2e` (host buffers in, host results out): the
database streams to the device in (tapered) row chunks overlapped with K1/K2, per-chunk lists are merged by K3.  Its
results must equal the device-resident pipeline bit for bit, for every arm and any chunking (retrieval/evaluator.py and
utils/cv_evaluator.py score each query against the whole database at once; chunking is ours)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from emr2a_b200.engine import get_engine
    return get_engine()


def _case(seed, n, q, d0, d1, c=3):
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, c, n).astype(np.int32)
    ql = rng.integers(0, c, q).astype(np.int32)
    cen0, cen1 = rng.standard_normal((c, d0)).astype(np.float32), rng.standard_normal((c, d1)).astype(np.float32)
    mk = lambda cen, l, d: (cen[l] * 0.6 + rng.standard_normal((len(l), d))).astype(np.float32)
    return (mk(cen0, lab, d0), mk(cen1, lab, d1)), (mk(cen0, ql, d0), mk(cen1, ql, d1)), lab, ql


@pytest.mark.parametrize("prec,n,q,chunk", [("rescore", 60000, 700, 9000), ("rescore", 60000, 700, None),
                                            ("bf16x3", 30000, 300, 7000), ("fp32", 5000, 64, 1100)])
def test_host_path_equals_device_path(eng, prec, n, q, chunk):
    import torch
    from emr2a_b200 import native
    db, qs, lab, ql = _case(n + q, n, q, 128, 64)
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    k, c = 5, 3
    want = eng.search_and_vote(db, qs, lab, ql, c, k, db_flags=flags, q_flags=flags, k_list=[1, 3, 5], precision=prec)
    pin = lambda x: torch.from_numpy(x).pin_memory()
    got = eng.search_and_vote_host(tuple(pin(x) for x in db), tuple(pin(x) for x in qs), pin(lab), pin(ql), c, k,
                                   db_flags=flags, q_flags=flags, k_list=[1, 3, 5], precision=prec, chunk_rows=chunk)
    for name in ("top_idx", "top_scores", "top_labels", "pred_top1", "pred_vote", "pred_weighted", "hit_counts",
                 "vote_counts", "confusion"):
        assert torch.equal(got[name], want[name].cpu()), name
    row_bytes = (128 + 64) * 4
    assert got["h2d_bytes"] == n * row_bytes + q * row_bytes + (n + q) * 4
    assert got["precision"] == prec


def test_host_path_row_offset_and_reduce_fn(eng):
    """Two 'ranks' on one GPU: each searches its half of a host-resident database with its row offset, the lists are
    merged by reduce_fn -- the multi-GPU e2e shape of bench.py, with UNEQUAL halves (H2D-weighted shards)."""
    import torch
    from emr2a_b200 import native
    n, q, k, c = 40000, 400, 10, 3
    db, qs, lab, ql = _case(9, n, q, 96, 32)
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    want = eng.search_and_vote(db, qs, lab, ql, c, k, db_flags=flags, q_flags=flags, precision="rescore")
    cut = 14848                                             # 58 tiles of 256 rows: a 37 % / 63 % split
    halves = []
    for lo, hi in ((0, cut), (cut, n)):
        keep = {}

        def grab(keys, keep=keep):
            keep["keys"] = keys
            return keys
        eng.search_and_vote_host((db[0][lo:hi], db[1][lo:hi]), qs, lab, ql, c, k, db_flags=flags, q_flags=flags,
                                 precision="rescore", row_offset=lo, reduce_fn=grab, chunk_rows=5000)
        halves.append(keep["keys"])
    merged = eng.topk_merge(torch.stack(halves), k)
    assert torch.equal(merged, want["keys"])
