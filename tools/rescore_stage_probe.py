"""Per-stage CUDA-event times of the rescore arm (filter / re-scoring / exact re-scan of 8 queries) with materialised and
with deferred fp32 database rows, on the C2 and C4 shapes."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine
eng = get_engine(); dev = eng.device
flags = native.NF_SEGNORM | native.NF_ROWNORM
def ev(): return torch.cuda.Event(enable_timing=True)
for name, n, d0, d1, dt in (("c2", 1_000_000, 512, 512, None), ("c4", 2_000_000, 4096, 1024, torch.bfloat16)):
    q, k, c = 10_000, 10, 3
    di = synth.device_block(0, n, d0, c, 17, dev, label_seed=17, dtype=dt)[0]
    dj = synth.device_block(0, n, d1, c, 18, dev, label_seed=17, dtype=dt)[0]
    qi = synth.device_block(50_003_968, q, d0, c, 17, dev, label_seed=17, dtype=dt)[0]
    qj = synth.device_block(50_003_968, q, d1, c, 18, dev, label_seed=17, dtype=dt)[0]
    qs = eng.prepare(qi, qj, 1.0, 1.0, flags, "rescore")
    flagged = torch.arange(0, 8, dtype=torch.int32, device=dev)
    for defer in (False, True):
        e = [ev() for _ in range(5)]
        for it in range(3):
            e[0].record(); db = eng.prepare(di, dj, 1.0, 1.0, flags, "rescore", defer_f32=defer)
            e[1].record(); cand, tau, kth = eng.topk_filter(qs, db, k)
            e[2].record(); pay = eng.rescore_candidates(cand, tau, None, qs, db, k)
            e[3].record(); comp = eng.exact_rescan(qs, db, flagged, k)
            e[4].record(); torch.cuda.synchronize()
            if it < 2:
                del db, cand, tau, kth, pay, comp
        t = [e[i].elapsed_time(e[i + 1]) for i in range(4)]
        print(f"{name} deferred={defer}: K1 db {t[0]:.3f} ms  filter {t[1]:.3f} ms  re-scoring {t[2]:.3f} ms  exact re-scan of 8 queries {t[3]:.3f} ms "
              f"(one pass over the rows: {n * (d0 + d1) * (2 if (defer and dt is not None) else 4) / t[3] / 1e9:.2f} TB/s)", flush=True)
        del db, cand, tau, kth, pay, comp
    del di, dj, qi, qj, qs
    torch.cuda.empty_cache()
