"""Oracle-sample parity at BASELINE.json's FULL sizes (C2, C3, C4 and a 10M-row C5 case on one GPU).

The database stays on the GPU; a sample of queries is run through the reference's own loop on the host
(``oracle.StreamedReferenceSample``: normalise/fuse per chunk, sgemv / sgemm per chunk, then the reference's
``np.argsort(scores)[-k:][::-1]`` over ALL N scores and the python votes -- utils/cv_evaluator.py:232-237 (late),
:269-300 (concat), :349-376 (fold rule)).  The bar: scores within 1e-5, Top-K index rows, majority and weighted votes
identical wherever adjacent score gaps exceed 2e-5 (the gap rule), for every sampled query."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAMPLE = 32
TOL = 1e-5


@pytest.fixture(scope="module")
def eng():
    from emr2a_b200.engine import get_engine
    return get_engine()


@pytest.fixture(scope="module")
def feeder():
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import streamed_feed
    return streamed_feed


def _free():
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()


def _check(oracle, ref, res, pick, name):
    got_idx = res["top_idx"][pick].cpu().numpy()
    got_sc = res["top_scores"][pick].cpu().numpy()
    par = oracle.sample_parity(ref, got_idx, got_sc, res["pred_vote"][pick].cpu().numpy(),
                               res["pred_weighted"][pick].cpu().numpy(), tol=TOL)
    print(f"{name}: {par}")
    assert par["ok"], (name, par)
    assert par["clear_rows"] >= int(0.7 * len(pick)), (name, par)       # the test must actually bite
    assert np.array_equal(res["pred_top1"][pick].cpu().numpy()[ref["top_idx"][:, 0] == got_idx[:, 0]],
                          ref["pred_top1"][ref["top_idx"][:, 0] == got_idx[:, 0]])
    return par


def test_c2_full_size_oracle_sample(eng, oracle, feeder):
    """C2: 1M x (512+512) fp32, concat fusion, 10k queries, K=10 (BASELINE.json configs[1])."""
    import torch
    from emr2a_b200 import native, synth
    _free()
    dev = eng.device
    n, d, n_q, k, c, seed = 1_000_000, 512, 10_000, 10, 3, 11
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    di, _ = synth.device_block(0, n, d, c, seed, dev, label_seed=seed)
    dt, _ = synth.device_block(0, n, d, c, seed + 1, dev, label_seed=seed)
    qi, ql = synth.device_block(50_003_968, n_q, d, c, seed, dev, label_seed=seed)
    qt, _ = synth.device_block(50_003_968, n_q, d, c, seed + 1, dev, label_seed=seed)
    labels = synth.device_labels(0, n, c, seed, dev)
    res = eng.search_and_vote((di, dt), (qi, qt), labels, ql, c, k, db_flags=flags, q_flags=flags, k_list=[1, 3, 5, k],
                              precision="rescore")
    assert res["precision"] == "rescore"
    pick = torch.linspace(0, n_q - 1, SAMPLE, device=dev).long()
    s = oracle.StreamedReferenceSample("concat", k, n, qi[pick].cpu().numpy(), qt[pick].cpu().numpy())
    feeder.feed(s, n, feeder.device_fetcher(di, dt), timed_rows=65536)
    ref = s.finish(labels.cpu().numpy(), ql[pick].cpu().numpy())
    _check(oracle, ref, res, pick, "c2")
    # K=5 (the other K of configs[1]) through the same operands: prefix property + its own oracle ranking
    res5 = eng.search_and_vote((di, dt), (qi, qt), labels, ql, c, 5, db_flags=flags, q_flags=flags, k_list=[1, 3, 5],
                               precision="rescore")
    assert torch.equal(res5["top_idx"], res["top_idx"][:, :5])
    s.k = 5
    _check(oracle, s.finish(labels.cpu().numpy(), ql[pick].cpu().numpy()), res5, pick, "c2 K=5")


def test_c3_full_size_oracle_sample(eng, oracle, feeder):
    """C3: late fusion, 5M cases, 512-d image + 512-d text, score = w*<T,T> + (1-w)*<I,I> over ALL rows, then Top-K
    (merge-then-Top-K, utils/cv_evaluator.py:232-237), w_text = 0.25, K=10."""
    import torch
    from emr2a_b200 import native, synth
    _free()
    dev = eng.device
    n, d, n_q, k, c, seed, w_text = 5_000_000, 512, 10_000, 10, 3, 13, 0.25
    di, _ = synth.device_block(0, n, d, c, seed, dev, label_seed=seed)
    dt, _ = synth.device_block(0, n, d, c, seed + 1, dev, label_seed=seed)
    qi, ql = synth.device_block(50_003_968, n_q, d, c, seed, dev, label_seed=seed)
    qt, _ = synth.device_block(50_003_968, n_q, d, c, seed + 1, dev, label_seed=seed)
    labels = synth.device_labels(0, n, c, seed, dev)
    res = eng.search_and_vote((di, dt), (qi, qt), labels, ql, c, k, db_flags=native.NF_SEGNORM, q_flags=native.NF_SEGNORM,
                              q_weights=(np.float32(1 - w_text), np.float32(w_text)), k_list=[1, 3, 5, k], precision="rescore")
    pick = torch.linspace(0, n_q - 1, SAMPLE, device=dev).long()
    s = oracle.StreamedReferenceSample("late", k, n, qi[pick].cpu().numpy(), qt[pick].cpu().numpy(), w_text=w_text)
    feeder.feed(s, n, feeder.device_fetcher(di, dt), timed_rows=65536)
    ref = s.finish(labels.cpu().numpy(), ql[pick].cpu().numpy())
    _check(oracle, ref, res, pick, "c3")


def test_c4_full_size_oracle_sample(eng, oracle, feeder):
    """C4: Qwen3-VL-shaped 4096-d image + 1024-d text, 2M cases, bf16 INPUTS with fp32 accumulation: the oracle consumes
    the same bf16 values widened to fp32 (exact), so the 1e-5 bar applies unchanged."""
    import torch
    from emr2a_b200 import native, synth
    _free()
    dev = eng.device
    n, d_img, d_txt, n_q, k, c, seed = 2_000_000, 4096, 1024, 10_000, 10, 3, 17
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    bf = torch.bfloat16
    di, _ = synth.device_block(0, n, d_img, c, seed, dev, label_seed=seed, dtype=bf)
    dt, _ = synth.device_block(0, n, d_txt, c, seed + 1, dev, label_seed=seed, dtype=bf)
    qi, ql = synth.device_block(50_003_968, n_q, d_img, c, seed, dev, label_seed=seed, dtype=bf)
    qt, _ = synth.device_block(50_003_968, n_q, d_txt, c, seed + 1, dev, label_seed=seed, dtype=bf)
    labels = synth.device_labels(0, n, c, seed, dev)
    res = eng.search_and_vote((di, dt), (qi, qt), labels, ql, c, k, db_flags=flags, q_flags=flags, k_list=[1, 3, 5, k],
                              precision="rescore")
    pick = torch.linspace(0, n_q - 1, SAMPLE, device=dev).long()
    s = oracle.StreamedReferenceSample("concat", k, n, qi[pick].float().cpu().numpy(), qt[pick].float().cpu().numpy())
    feeder.feed(s, n, feeder.device_fetcher(di, dt), chunk_rows=16384, timed_rows=16384)
    ref = s.finish(labels.cpu().numpy(), ql[pick].cpu().numpy())
    _check(oracle, ref, res, pick, "c4")


def test_c5_ten_million_rows_fold_rule_oracle_sample(eng, oracle, feeder):
    """C5 shape on one GPU: a 10M-case 1024-d fused database in fold order, 5 folds; 16k of the cases are queries, each
    searched against the cases of the OTHER four folds (K=5).  The oracle scores every sampled query against all 8M
    admissible rows."""
    import torch
    from emr2a_b200 import native, synth
    _free()
    dev = eng.device
    n, d, k, c, seed, n_folds = 10_000_000, 512, 5, 3, 19, 5
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    di, _ = synth.device_block(0, n, d, c, seed, dev, label_seed=seed)
    dt, _ = synth.device_block(0, n, d, c, seed + 1, dev, label_seed=seed)
    labels = synth.device_labels(0, n, c, seed, dev)
    fold = (torch.arange(n, device=dev, dtype=torch.int64) * n_folds // n).to(torch.uint8)
    db = eng.prepare(di, dt, 1.0, 1.0, flags, "rescore")
    q_rows = torch.arange(0, n, 610, device=dev)[:16384]                    # ascending => fold-sorted queries
    qs = eng.prepare(di[q_rows], dt[q_rows], 1.0, 1.0, flags, "rescore")
    keys = eng.topk_search(qs, db, k, "rescore", q_fold=fold[q_rows], db_fold=fold, fold_sorted=True)
    unverified, overflow = eng.consume_status()
    assert not overflow
    res = eng.vote_metrics(keys, labels, labels[q_rows], c, k_list=[1, 3, 5], q_group=fold[q_rows], n_groups=n_folds)
    assert not bool((fold[res["top_idx"]] == fold[q_rows][:, None]).any())
    pick = torch.linspace(0, len(q_rows) - 1, 24, device=dev).long()
    rows = q_rows[pick]
    s = oracle.StreamedReferenceSample("concat", k, n, di[rows].cpu().numpy(), dt[rows].cpu().numpy(),
                                       q_fold=fold[rows].cpu().numpy())
    feeder.feed(s, n, feeder.device_fetcher(di, dt, fold), timed_rows=65536)
    ref = s.finish(labels.cpu().numpy(), labels[rows].cpu().numpy())
    assert ref["rows_scored"] == 24 * (n - n // n_folds)
    _check(oracle, ref, res, pick, "c5 10M")
    del db, qs
    _free()
