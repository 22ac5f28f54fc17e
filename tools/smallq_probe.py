"""K2 alone at small Q (resident operands): time of emr2a_topk_search per precision arm / kernel variant."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine
eng = get_engine(); dev = eng.device
n, d, k, c = 1_000_000, 512, 10, 3
flags = native.NF_SEGNORM | native.NF_ROWNORM
di, _ = synth.device_block(0, n, d, c, 11, dev, label_seed=11); dt, _ = synth.device_block(0, n, d, c, 12, dev, label_seed=11)
for prec in os.environ.get("PRECS", "rescore,bf16x1").split(","):
    db = eng.prepare(di, dt, 1.0, 1.0, flags, prec)
    for q in [int(v) for v in os.environ.get("QS", "1,64,128,256,512").split(",")]:
        qi, _ = synth.device_block(50_003_968, q, d, c, 11, dev, label_seed=11); qt, _ = synth.device_block(50_003_968, q, d, c, 12, dev, label_seed=11)
        qs = eng.prepare(qi, qt, 1.0, 1.0, flags, prec)
        for _ in range(3):
            eng.topk_search(qs, db, k, prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.topk_search(qs, db, k, prec)
        e1.record(); torch.cuda.synchronize()
        eng.consume_status()
        ms = e0.elapsed_time(e1) / 20
        print(f"TC2={os.environ.get('EMR2A_TC2', '1')} [{prec}] Q={q:4d}: emr2a_topk_search {ms:.3f} ms -> plane stream {n*2*d*2/1e9/ms*1e3:.0f} GB/s", flush=True)
