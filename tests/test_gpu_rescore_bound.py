"""The rescore arm's verification bound at long contractions (emr2a_b200/csrc/rescore.cu: error_bound).

With rows that are exactly representable in bf16 the two quantisation terms of the bound vanish (r_q = r_d = 0) and
the verification rests entirely on the ARITHMETIC term, (1.02 D + 8) * 2^-23 * n_q * n_d -- the corner the round-1
bound (a constant 1e-5) did not cover.  D = 5120 (the Qwen3-VL-shaped C4 width)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

D = 5120


@pytest.fixture(scope="module")
def eng():
    from emr2a_b200.engine import get_engine
    return get_engine()


def _bf16_exact(x):
    """Round float32 values to the nearest bf16 (ties to even), returned as float32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def _ulp_down(x, n):
    """Move bf16-exact float32 values n bf16 ulps towards zero (stay bf16-exact)."""
    u = x.view(np.uint32).copy()
    return (u - np.uint32(n << 16)).view(np.float32)


def test_tensor_core_accumulation_error_is_far_inside_the_bound(eng):
    """Measure |s~ - s| of the 1-pass filter on bf16-exact operands (products exact, only the fp32 accumulation of the
    tensor pipe remains) against float64, D = 5120: it must sit below the arithmetic term of the bound."""
    import torch
    from emr2a_b200 import native
    rng = np.random.default_rng(3)
    Q, N, K = 256, 4096, 10
    q = _bf16_exact(rng.standard_normal((Q, D)).astype(np.float32) / np.sqrt(D))
    db = _bf16_exact(rng.standard_normal((N, D)).astype(np.float32) / np.sqrt(D))
    # worst-case-ish sign structure: all products positive for a quarter of the pairs (no cancellation, largest partial sums)
    q[:64] = np.abs(q[:64]); db[:1024] = np.abs(db[:1024])
    qo = eng.normalize_fuse(q, flags=0, want_f32=True, want_planes=True, want_lo=False)
    do = eng.normalize_fuse(db, flags=0, want_f32=True, want_planes=True, want_lo=False)
    ws_bytes = int(eng.lib.emr2a_topk_search_workspace_bytes(Q, N, D, K, native.PREC_BF16X1))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=eng.device)
    wsp = (ws.data_ptr() + 255) // 256 * 256
    keys = torch.zeros((Q, K), dtype=torch.int64, device=eng.device)
    dump = torch.full((Q, N), float("nan"), dtype=torch.float32, device=eng.device)
    native.check(eng.lib.emr2a_debug_topk_search_dump(qo.hi.data_ptr(), None, do.hi.data_ptr(), None, Q, N, D,
                                                      qo.hi.stride(0), do.hi.stride(0), None, None, 0, K,
                                                      native.PREC_BF16X1, keys.data_ptr(), wsp, ws_bytes, dump.data_ptr(), None))
    torch.cuda.synchronize()
    truth = q.astype(np.float64) @ db.astype(np.float64).T
    err = np.abs(dump.cpu().numpy().astype(np.float64) - truth)
    nq = np.linalg.norm(q.astype(np.float64), axis=1)[:, None]
    nd = np.linalg.norm(db.astype(np.float64), axis=1)[None, :]
    bound = (1.02 * D + 8) * 2.0 ** -23 * nq * nd
    rel = float((err / (nq * nd)).max())
    print(f"tensor-core fp32 accumulation, D={D}: max |err| {err.max():.3e}, max err/(n_q n_d) {rel:.3e}, "
          f"bound coefficient {(1.02 * D + 8) * 2.0 ** -23:.3e}")
    assert np.all(err <= bound)
    # measured on B200: 1.4e-5 * n_q n_d for the all-positive pairs (the accumulator truncates: ~115 units of 2^-23
    # over 320 accumulation steps) -- ABOVE the constant 1e-5 n_q n_d the round-1 bound used, 45x below the new term
    assert rel < 0.1 * (1.02 * D + 8) * 2.0 ** -23


def test_bf16_exact_rows_with_planted_near_ties(eng, oracle):
    """bf16-exact database and queries (r_q = r_d = 0), D = 5120, planted ladders of near-tied rows: row j of a ladder
    is the query with `step * j` coordinates moved one bf16 ulp towards zero, so its score sits j * ~gap below the
    query's self-score, all values exactly representable.  Ladders of 100 rows are longer than the 64 merged
    candidates, i.e. they straddle the candidate cut; gaps run from ~5e-6 to ~5e-5.  Whatever the verification
    decides (verified, or exact re-scan), the result must be the exact Top-K: index rows identical to a float64
    ranking wherever the float64 gaps are clear of the fp32 rounding of the re-scoring (the fp32 arm is held to the
    same rule), scores within 1e-5."""
    import torch
    from emr2a_b200.engine import unpack_keys
    rng = np.random.default_rng(11)
    N, Q, K = 30_000, 48, 10
    db = _bf16_exact(rng.standard_normal((N, D)).astype(np.float32) / np.sqrt(D))
    qs = _bf16_exact(rng.standard_normal((Q, D)).astype(np.float32) / np.sqrt(D))
    # one bf16 ulp on coordinate i moves the score by d_i = |q_i| * ulp(q_i) (~1e-6 for the typical |q_i| ~ 1/72); a rung
    # collects coordinates until the score has dropped by the ladder's gap
    ladders = {2: 5e-6, 9: 1e-5, 17: 2.5e-5, 30: 5e-5, 41: 7e-6}              # query -> target gap between rungs
    for h, (qi, gap) in enumerate(ladders.items()):
        q = qs[qi]
        d = np.abs(q).astype(np.float64) * np.abs(q - _ulp_down(q, 1)).astype(np.float64)
        pool = [i for i in rng.permutation(D) if 0 < d[i] <= gap / 3]
        row, used, rungs = q.copy(), 0, 0
        db[2000 * (h + 1)] = row                                              # rung 0: the query itself
        for j in range(1, 100):
            drop = 0.0
            while drop < gap and used < len(pool):
                i = pool[used]; used += 1
                row[i] = _ulp_down(row[i:i + 1], 1)[0]
                drop += d[i]
            if drop < gap:
                break
            db[2000 * (h + 1) + 13 * j] = row
            rungs = j
        assert rungs >= 70, (qi, rungs)                                       # longer than the 64 merged candidates
    assert np.array_equal(db, _bf16_exact(db)) and np.array_equal(qs, _bf16_exact(qs))
    truth = qs.astype(np.float64) @ db.astype(np.float64).T
    order = np.argsort(-truth, axis=1, kind="stable")
    want_idx = order[:, :K]
    ladder_sc = np.take_along_axis(truth, order[:, :K + 1], axis=1)
    gaps = -np.diff(ladder_sc, axis=1)
    assert 4e-6 < gaps[2].min() < 8e-6 and 4e-5 < gaps[30].min() < 7e-5        # the planted ladders are what they claim
    # fp32 re-scoring error at D = 5120 is ~1e-7 (lane-strided FMAs); rows whose float64 gaps exceed 2e-6 must be exact
    clear = gaps.min(axis=1) > 2e-6
    assert clear[list(ladders)].all() and clear.mean() > 0.9
    for prec in ("rescore", "fp32"):
        qo = eng.prepare(qs, flags=0, precision=prec)
        do = eng.prepare(db, flags=0, precision=prec)
        if prec == "rescore":
            st = do.stats.cpu().numpy()
            assert st[1] == 0.0 and qo.stats.cpu().numpy()[1] == 0.0          # r_d = r_q = 0: bf16-exact rows
        keys = eng.topk_search(qo, do, K, prec)
        if prec == "rescore":
            unverified, overflow = eng.consume_status()
            flagged = eng.last_unverified.cpu().numpy().astype(bool)
            assert not overflow
            # the fine ladders (gap ~5e-6: 10th best to 64th candidate closer than the arithmetic bound) cannot verify
            assert flagged[2] and flagged[41], (unverified, np.flatnonzero(flagged))
            # random rows at D = 5120 with r = 0: the bound is ~6e-4, scores of unrelated rows are ~N(0, 1/72): verified
            assert flagged.sum() <= len(ladders) + 2, np.flatnonzero(flagged)
        sc, idx = unpack_keys(keys)
        assert np.max(np.abs(sc - np.take_along_axis(truth, want_idx, axis=1))[clear]) < 1e-5
        assert np.array_equal(idx[clear], want_idx[clear]), prec
        par = oracle.sample_parity({"top_idx": want_idx, "top_scores": np.take_along_axis(truth, want_idx, axis=1),
                                    "next_score": ladder_sc[:, K]}, idx, sc, tol=1e-5)
        assert par["ok"], (prec, par)
