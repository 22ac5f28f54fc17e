"""Parity at BASELINE.json's full sizes through size-independent properties (the CPU oracle would need hours):
C2 (1M x (512+512), 10k queries, K=10) and a C5-shaped all-queries fold-masked search."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from emr2a_b200.engine import get_engine
    return get_engine()


def test_c2_full_size_properties(eng):
    import torch
    from emr2a_b200 import native, synth
    from emr2a_b200.engine import unpack_keys
    dev = eng.device
    n, d, n_q, k, c = 1_000_000, 512, 10_000, 10, 3
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    di, _ = synth.device_block(0, n, d, c, 11, dev, label_seed=11)
    dt, _ = synth.device_block(0, n, d, c, 12, dev, label_seed=11)
    qi, ql = synth.device_block(50_003_968, n_q, d, c, 11, dev, label_seed=11)
    qt, _ = synth.device_block(50_003_968, n_q, d, c, 12, dev, label_seed=11)
    labels = synth.device_labels(0, n, c, 11, dev)
    # planted queries: the first 512 queries are database rows (scaled: cosine ignores the length)
    planted = torch.arange(0, 512, device=dev) * 1931 + 17
    qi[:512] = di[planted] * 3.0
    qt[:512] = dt[planted] * 0.5
    ql[:512] = labels[planted]

    db_r = eng.prepare(di, dt, 1.0, 1.0, flags, "rescore")
    q_r = eng.prepare(qi, qt, 1.0, 1.0, flags, "rescore")
    keys = eng.topk_search(q_r, db_r, k, "rescore")
    unverified, overflow = eng.consume_status()
    assert not overflow and unverified == 0
    sc, idx = unpack_keys(keys)
    # (1) self retrieval: Top-1 of a planted query is its own row, cosine 1
    assert np.array_equal(idx[:512, 0], planted.cpu().numpy())
    assert np.max(np.abs(sc[:512, 0] - 1.0)) < 2e-6
    # (2) lists are sorted (score desc, index asc), indices valid and distinct
    assert np.all(np.diff(sc, axis=1) <= 0) and idx.min() >= 0 and idx.max() < n
    assert all(len(set(row)) == k for row in idx[::97])
    # (3) union property: Top-K over two row ranges, merged, is bit-identical to the Top-K over all rows
    half = 499_968                                                # 256-row aligned, as dist.shard_range cuts
    parts = []
    for lo, hi in ((0, half), (half, n)):
        parts.append(eng.topk_search(q_r, eng._rows(db_r, lo, hi), k, "rescore", idx_base=lo))
    assert eng.consume_status() == (0, False)
    merged = eng.topk_merge(torch.stack(parts), k)
    assert torch.equal(merged, keys)
    # (4) the 3-pass tensor-core arm agrees: scores to 1e-5, identical rows where the gaps are clear
    db_3 = eng.prepare(di, dt, 1.0, 1.0, flags, "bf16x3")
    q_3 = eng.prepare(qi, qt, 1.0, 1.0, flags, "bf16x3")
    s3, i3 = unpack_keys(eng.topk_search(q_3, db_3, k, "bf16x3"))
    assert np.max(np.abs(s3 - sc)) < 1e-5
    clear = np.abs(np.diff(sc, axis=1)).min(axis=1) > 2e-5
    assert clear.mean() > 0.5 and np.array_equal(i3[clear], idx[clear])
    del db_3, q_3
    # (5) exact fp32 scores: re-compute the winners' cosines in float64 from the raw rows
    pick = torch.arange(0, n_q, 250, device=dev)
    rows = torch.from_numpy(idx).to(dev)[pick]                    # [40, k]
    def unit(x):
        x = x.double()
        return x / x.norm(dim=-1, keepdim=True)
    qa = torch.cat([unit(qi[pick]), unit(qt[pick])], dim=1) / np.sqrt(2.0)
    da = torch.cat([unit(di[rows]), unit(dt[rows])], dim=2) / np.sqrt(2.0)
    want = torch.einsum("qd,qkd->qk", qa, da).cpu().numpy()
    assert np.max(np.abs(want - sc[pick.cpu().numpy()])) < 1e-5
    # (6) vote counters are consistent with the lists
    r = eng.vote_metrics(keys, labels, ql, c, k_list=[1, 3, 5, k])
    hits = r["hit_counts"][0].cpu().numpy()
    assert hits[0] <= hits[1] <= hits[2] <= hits[3] <= n_q and hits[0] >= 512
    assert int(r["confusion"][0].sum()) == 2 * n_q and int(r["group_sizes"][0]) == n_q
    top_lab = labels[torch.from_numpy(idx).to(dev)]
    assert torch.equal(r["pred_top1"], top_lab[:, 0])
    assert int((top_lab[:, 0] == ql).sum()) == int(r["vote_counts"][0, 0]) == hits[0]


def test_c5_shaped_all_queries_fold_rule(eng):
    """Every row a query against the rows of the other folds (fold-sorted, K=5), 1M x 1024: no neighbour shares the
    query's fold, a duplicated row in another fold is found with cosine 1, and the result does not depend on the
    query block size."""
    import torch
    from emr2a_b200 import native, synth
    dev = eng.device
    n, d, c, k, n_folds = 1_000_000, 1024, 3, 5, 5
    x, lab = synth.device_block(0, n, d, c, 19, dev)
    folds = (torch.arange(n, device=dev) * n_folds // n).to(torch.uint8)          # sorted, equal sizes
    x[700_000:700_100] = x[100_000:100_100]                                       # fold 3 copies of fold 0 rows
    out = eng.cv_search_and_vote((x,), lab, folds, c, k, flags=native.NF_ROWNORM, k_list=[1, 3, 5], n_folds=n_folds,
                                 q_block=1 << 18)
    assert out["unverified"] == 0
    idx = out["top_idx"]
    assert int(idx.min()) >= 0
    assert not bool((folds[idx] == folds[:, None]).any())
    assert torch.equal(idx[700_000:700_100, 0], torch.arange(100_000, 100_100, device=dev))
    assert torch.equal(idx[100_000:100_100, 0], torch.arange(700_000, 700_100, device=dev))
    assert float((out["top_scores"][700_000:700_100, 0] - 1.0).abs().max()) < 2e-6
    assert int(out["group_sizes"].sum()) == n and int(out["hit_counts"][:, 0].sum()) == int((lab[idx[:, 0]] == lab).sum())
    again = eng.cv_search_and_vote((x,), lab, folds, c, k, flags=native.NF_ROWNORM, k_list=[1, 3, 5], n_folds=n_folds,
                                   q_block=1 << 20)
    assert torch.equal(again["top_idx"], idx) and torch.equal(again["hit_counts"], out["hit_counts"])
    assert torch.equal(again["confusion"], out["confusion"])


def test_sharded_cv_single_rank_equals_engine_path(eng):
    """dist.sharded_cv_search_and_vote with one rank (no process group) is the engine's one-pass CV; unsorted fold
    vectors use the per-element mask only.  (World sizes 2 and 8: tools/dist_check.py under torchrun.)"""
    import torch
    from emr2a_b200 import native, synth
    from emr2a_b200.dist import sharded_cv_search_and_vote
    dev = eng.device
    n, d, c, k = 60_000, 192, 3, 5
    xi, lab = synth.device_block(0, n, d, c, 23, dev)
    xt, _ = synth.device_block(0, n, d, c, 24, dev, label_seed=23)
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    keys = ("hit_counts", "vote_counts", "confusion", "group_sizes", "top_idx", "top_scores", "pred_vote", "pred_weighted")
    sorted_folds = (torch.arange(n, device=dev) * 5 // n).to(torch.uint8)
    a = sharded_cv_search_and_vote(eng, (xi, xt), lab, sorted_folds, c, k, 0, flags, precision="rescore", q_block=16384,
                                   want_lists=True)
    b = eng.cv_search_and_vote((xi, xt), lab, sorted_folds, c, k, flags=flags, precision="rescore", n_folds=5)
    assert all(torch.equal(a[key], b[key]) for key in keys)
    mixed = (torch.arange(n, device=dev) % 5).to(torch.uint8)
    a = sharded_cv_search_and_vote(eng, (xi, xt), lab, mixed, c, k, 0, flags, precision="rescore", q_block=16384,
                                   fold_sorted=False, want_lists=True)
    b = eng.cv_search_and_vote((xi, xt), lab, mixed, c, k, flags=flags, precision="rescore", n_folds=5)
    assert all(torch.equal(a[key], b[key]) for key in keys)
    assert not bool((mixed[a["top_idx"]] == mixed[:, None]).any())
