"""Small-batch serving latency against a resident index: Q queries vs a 1M x 1024 database (K = 10).
At small Q the search is HBM-bound (the bf16 database plane, 2.05 GB, must be streamed once per batch)."""
import os, sys, time, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine
eng = get_engine(); dev = eng.device
n, d, k, c = int(os.environ.get("N", 1_000_000)), 512, 10, 3
flags = native.NF_SEGNORM | native.NF_ROWNORM
di, _ = synth.device_block(0, n, d, c, 11, dev, label_seed=11); dt, _ = synth.device_block(0, n, d, c, 12, dev, label_seed=11)
labels = synth.device_labels(0, n, c, 11, dev)
for prec in os.environ.get("PRECS", "rescore,bf16x3").split(","):
    index = eng.build_index((di, dt), labels, c, flags=flags, precision=prec, k=k)
    for q in [int(v) for v in os.environ.get("QS", "1,8,64,128,256,1024,4096").split(",")]:
        qi, ql = synth.device_block(50_003_968, q, d, c, 11, dev, label_seed=11); qt, _ = synth.device_block(50_003_968, q, d, c, 12, dev, label_seed=11)
        for _ in range(3):
            r = index.search((qi, qt), ql, k=k, want_lists=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            r = index.search((qi, qt), ql, k=k, want_lists=False)
        e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        ms = e0.elapsed_time(e1) / reps
        g_ms = None
        if q <= 1024:                                  # the same batch as one CUDA graph
            graphed = index.capture(q, k=k, want_lists=False, seg_dims=[d, d])
            for _ in range(3):
                graphed((qi, qt), ql)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                graphed((qi, qt), ql)
            torch.cuda.synchronize()
            g_ms = (time.perf_counter() - t0) / reps * 1e3
            del graphed
        plane_gb = n * 2 * d * 2 / 1e9 * (2 if prec == "bf16x3" else 1)
        print(f"[{prec}] Q={q:5d}: {ms:.3f} ms/batch (wall {wall:.3f}) = {q/ms*1e3:9.0f} queries/s; database plane stream "
              f"{plane_gb/ms*1e3:.0f} GB/s; {2*q*n*2*d/ms/1e9:.0f} TFLOP/s" + (f"; CUDA graph {g_ms:.3f} ms wall" if g_ms else ""), flush=True)
    del index
