import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine
eng = get_engine(); dev = eng.device
n, q, k, c = 2_000_000, 10_000, 10, 3
flags = native.NF_SEGNORM | native.NF_ROWNORM
di = synth.device_block(0, n, 4096, c, 17, dev, label_seed=17)[0].to(torch.bfloat16)
dt = synth.device_block(0, n, 1024, c, 18, dev, label_seed=17)[0].to(torch.bfloat16)
qi = synth.device_block(50_003_968, q, 4096, c, 17, dev, label_seed=17)[0].to(torch.bfloat16)
qt = synth.device_block(50_003_968, q, 1024, c, 18, dev, label_seed=17)[0].to(torch.bfloat16)
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(3):
    e = [ev() for _ in range(4)]
    e[0].record(); db = eng.prepare(di, dt, 1.0, 1.0, flags, "rescore")
    e[1].record(); qs = eng.prepare(qi, qt, 1.0, 1.0, flags, "rescore")
    e[2].record(); keys = eng.topk_search(qs, db, k, "rescore")
    e[3].record(); torch.cuda.synchronize(); st = eng.consume_status()
    print(f"iter {it}: K1 db {e[0].elapsed_time(e[1]):.2f} ms (82 GB -> {82.0/e[0].elapsed_time(e[1]):.2f} TB/s)  K1 q {e[1].elapsed_time(e[2]):.2f}  K2 {e[2].elapsed_time(e[3]):.2f} ms ({2*5120*q*n/e[2].elapsed_time(e[3])/1e9:.0f} TFLOP/s)  {st}", flush=True)
    del db, qs, keys
