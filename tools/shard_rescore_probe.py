"""The re-scoring stage of ONE rank of an 8-GPU cooperative search, on a single GPU: filter the whole 1M-row database for
the global K-th best filter score (what the all-reduce delivers), then time emr2a_rescore_candidates on shard 0
(125k rows) with that floor -- materialised and deferred fp32 rows."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine, Operand
eng = get_engine(); dev = eng.device
flags = native.NF_SEGNORM | native.NF_ROWNORM
n, d, q, k, c, parts = 1_000_000, 512, 10_000, 10, 3, int(os.environ.get("PARTS", 8))
di = synth.device_block(0, n, d, c, 1234, dev, label_seed=1234)[0]
dj = synth.device_block(0, n, d, c, 1235, dev, label_seed=1234)[0]
qi = synth.device_block(50_003_968, q, d, c, 1234, dev, label_seed=1234)[0]
qj = synth.device_block(50_003_968, q, d, c, 1235, dev, label_seed=1234)[0]
qs = eng.prepare(qi, qj, 1.0, 1.0, flags, "rescore")
full = eng.prepare(di, dj, 1.0, 1.0, flags, "rescore", defer_f32=True)
_, _, floor = eng.topk_filter(qs, full, k)
del full
hi = n // parts
for defer in (False, True):
    db = eng.prepare(di[:hi], dj[:hi], 1.0, 1.0, flags, "rescore", defer_f32=defer)
    cand, tau, kth = eng.topk_filter(qs, db, k)
    for _ in range(3):
        pay = eng.rescore_candidates(cand, tau, floor, qs, db, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pay = eng.rescore_candidates(cand, tau, floor, qs, db, k)
    e1.record(); torch.cuda.synchronize()
    keys, bounds = eng.split_payload(pay, q, k)
    print(f"shard 0 of {parts} ({hi} rows) deferred={defer}: re-scoring {e0.elapsed_time(e1) / 20:.3f} ms per call; "
          f"non-empty exact keys per query {float((keys != 0).sum()) / q:.2f}", flush=True)
