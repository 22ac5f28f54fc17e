"""Per-stage CUDA-event timing of one bench step (C2) to see where the non-K2 time goes."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine
eng = get_engine(); dev = eng.device
n, d, q, k, c = int(os.environ.get("N", 1_000_000)), 512, int(os.environ.get("Q", 10_000)), 10, 3
flags = native.NF_SEGNORM | native.NF_ROWNORM
di, _ = synth.device_block(0, n, d, c, 11, dev, label_seed=11); dt, _ = synth.device_block(0, n, d, c, 12, dev, label_seed=11)
qi, ql = synth.device_block(50_003_968, q, d, c, 11, dev, label_seed=11); qt, _ = synth.device_block(50_003_968, q, d, c, 12, dev, label_seed=11)
labels = synth.device_labels(0, n, c, 11, dev)
prec = os.environ.get("PREC", "rescore")
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(6):
    e = [ev() for _ in range(6)]
    e[0].record(); db = eng.prepare(di, dt, 1.0, 1.0, flags, prec, defer_f32=True)   # EMR2A_DEFER_F32=0: materialised fp32 rows
    e[1].record(); qs = eng.prepare(qi, qt, 1.0, 1.0, flags, prec)
    e[2].record(); keys = eng.topk_search(qs, db, k, prec)
    e[3].record(); st = eng.consume_status()
    e[4].record(); r = eng.vote_metrics(keys, labels, ql, c, k_list=[1, 3, 5, k])
    e[5].record(); torch.cuda.synchronize()
    names = ["K1 db", "K1 q", "K2 search", "status sync", "K4 vote"]
    print(f"iter {it} [{prec}] " + "  ".join(f"{nm} {e[i].elapsed_time(e[i+1]):.3f} ms" for i, nm in enumerate(names)) + f"  unverified {st}", flush=True)
