"""Time the K2 kernel alone (CUDA events) for several precision arms / K / shapes."""
import os, sys, json
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine

def main():
    eng = get_engine()
    dev = eng.device
    cfgs = [tuple(x) for x in json.loads(os.environ.get("SWEEP", '[[1000000,1024,10000]]'))]
    arms = os.environ.get("ARMS", "bf16x3:10,rescore:10,rescore:5,bf16x1:10").split(",")
    for (n, d, q) in cfgs:
        db, _ = synth.device_block(0, n, d, 3, 11, dev)
        qs, _ = synth.device_block(50_003_968, q, d, 3, 11, dev)
        for arm in arms:
            prec, k = arm.split(":"); k = int(k)
            dbo = eng.prepare(db, flags=native.NF_ROWNORM, precision=prec)
            qo = eng.prepare(qs, flags=native.NF_ROWNORM, precision=prec)
            for _ in range(2):
                eng.topk_search(qo, dbo, k, prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                eng.topk_search(qo, dbo, k, prec)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            tf = 2.0 * n * d * q / ms / 1e9
            print(f"N={n} D={d} Q={q} {prec} K={k}: {ms:.3f} ms  {tf:.1f} TFLOP/s algorithmic  {q/ms*1e3:.0f} q/s", flush=True)
            del dbo, qo
main()
