"""Is a single-query search slow because of the out-of-bounds rows of its query tile?  Compare Q=1 with a 64-row batch
that holds the same query 64 times, and with one real query + 63 in-bounds zero rows."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine
eng = get_engine(); dev = eng.device
n, d, k, c = 1_000_000, 512, 10, 3
flags = native.NF_SEGNORM | native.NF_ROWNORM
di, _ = synth.device_block(0, n, d, c, 11, dev, label_seed=11); dt, _ = synth.device_block(0, n, d, c, 12, dev, label_seed=11)
prec = "bf16x1"
db = eng.prepare(di, dt, 1.0, 1.0, flags, prec)
qi, _ = synth.device_block(50_003_968, 64, d, c, 11, dev, label_seed=11); qt, _ = synth.device_block(50_003_968, 64, d, c, 12, dev, label_seed=11)
cases = {"Q=1": (qi[:1], qt[:1]), "Q=64 distinct": (qi, qt), "Q=64 = one query 64x": (qi[:1].repeat(64, 1), qt[:1].repeat(64, 1))}
zi, zt = torch.zeros_like(qi), torch.zeros_like(qt); zi[0], zt[0] = qi[0], qt[0]
cases["Q=64 = one query + 63 zero rows"] = (zi, zt)
cases["Q=2"] = (qi[:2], qt[:2]); cases["Q=48"] = (qi[:48], qt[:48])
for rep in range(2):
    for name, (a, b) in cases.items():
        qs = eng.prepare(a.contiguous(), b.contiguous(), 1.0, 1.0, flags, prec)
        for _ in range(3):
            eng.topk_search(qs, db, k, prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.topk_search(qs, db, k, prec)
        e1.record(); torch.cuda.synchronize()
        print(f"{name:34s}: {e0.elapsed_time(e1)/20:.3f} ms", flush=True)
