"""Bring-up diagnostics for the tcgen05 kernel (run on the GPU box):
dumps the scores the epilogue saw and compares them with fp64 math, then the Top-K."""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))
from emr2a_b200 import native  # noqa: E402
from emr2a_b200.engine import get_engine, unpack_keys, _round_up  # noqa: E402
import emr2a_oracle as oracle  # noqa: E402


def run(Q, N, D, K, prec, fold=False, seed=0):
    eng = get_engine()
    rng = np.random.default_rng(seed)
    q = oracle.unit_rows(rng.standard_normal((Q, D)).astype(np.float32))
    db = oracle.unit_rows(rng.standard_normal((N, D)).astype(np.float32))
    qo = eng.normalize_fuse(q, flags=0, want_f32=True, want_planes=True)
    do = eng.normalize_fuse(db, flags=0, want_f32=True, want_planes=True)
    code = {"bf16x3": native.PREC_BF16X3, "bf16x1": native.PREC_BF16X1}[prec]
    ws_bytes = int(eng.lib.emr2a_topk_search_workspace_bytes(Q, N, D, K, code))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=eng.device)
    wsp = _round_up(ws.data_ptr(), 256)
    keys = torch.zeros((Q, K), dtype=torch.int64, device=eng.device)
    dump = torch.full((Q, N), float("nan"), dtype=torch.float32, device=eng.device)
    qf = dbf = None
    if fold:
        qf = torch.from_numpy(rng.integers(0, 5, Q).astype(np.uint8)).to(eng.device)
        dbf = torch.from_numpy(rng.integers(0, 5, N).astype(np.uint8)).to(eng.device)
    torch.cuda.synchronize()
    t0 = time.time()
    rc = eng.lib.emr2a_debug_topk_search_dump(qo.hi.data_ptr(), qo.lo.data_ptr(), do.hi.data_ptr(), do.lo.data_ptr(),
                                              Q, N, D, qo.hi.stride(0), do.hi.stride(0), native.ptr(qf), native.ptr(dbf),
                                              0, K, code, keys.data_ptr(), wsp, ws_bytes, dump.data_ptr(), None)
    if rc != 0:
        print("  rc", rc, native.last_error())
        return False
    torch.cuda.synchronize()
    dt = time.time() - t0
    got = dump.cpu().numpy()
    if prec == "bf16x1":
        qh = (qo.hi.cpu().numpy().view(np.uint16).astype(np.uint32) << 16).view(np.float32)[:, :D]
        dh = (do.hi.cpu().numpy().view(np.uint16).astype(np.uint32) << 16).view(np.float32)[:, :D]
        truth = qh.astype(np.float64) @ dh.astype(np.float64).T
    else:
        truth = q.astype(np.float64) @ db.astype(np.float64).T
    err = np.abs(got - truth)
    nan = int(np.isnan(got).sum())
    print(f"  {prec} Q={Q} N={N} D={D} K={K} fold={fold}: score max err {np.nanmax(err):.3e} nan {nan} ({dt*1e3:.1f} ms)")
    # top-k check against the dumped scores themselves (exact) and the fp64 truth (gap rule)
    sc, idx = unpack_keys(keys)
    ref = got.copy()
    if fold:
        ref = np.where(qf.cpu().numpy()[:, None] == dbf.cpu().numpy()[None, :], -np.inf, ref)
    order = np.argsort(-ref, axis=1, kind="stable")[:, :K]
    exact = np.array_equal(order, idx)
    print(f"     top-k from dumped scores exact: {exact}")
    if not exact:
        bad = np.nonzero((order != idx).any(axis=1))[0]
        print("     first bad rows", bad[:5], order[bad[0]], idx[bad[0]], sc[bad[0]], ref[bad[0]][order[bad[0]]])
    return nan == 0 and np.nanmax(err) < (3e-3 if prec == "bf16x1" and False else 1e-5) and exact


if __name__ == "__main__":
    ok = True
    for args in [(128, 256, 64, 5, "bf16x1"), (128, 256, 64, 5, "bf16x3"), (100, 1000, 128, 5, "bf16x3"),
                 (300, 5000, 200, 10, "bf16x3"), (300, 5000, 200, 10, "bf16x1"), (1000, 70000, 1024, 10, "bf16x3"),
                 (257, 3333, 96, 20, "bf16x3")]:
        ok = run(*args) and ok
    ok = run(300, 5000, 192, 5, "bf16x3", fold=True) and ok
    print("BRINGUP", "OK" if ok else "FAILED")
