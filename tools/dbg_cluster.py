import os, sys, torch, numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native
from emr2a_b200.engine import get_engine, unpack_keys
eng = get_engine()
g = torch.Generator(device="cuda").manual_seed(5)
D, Q, K = 256, 1500, 10
for n_clusters, per, noise in ((2500, 40, 0.05), (2500, 40, 0.2), (400, 250, 0.02)):
    centres = torch.randn((n_clusters, D), generator=g, device="cuda")
    db = centres.repeat_interleave(per, dim=0) + noise * torch.randn((n_clusters * per, D), generator=g, device="cuda")
    pick = torch.randint(0, n_clusters, (Q,), generator=g, device="cuda")
    qs = centres[pick] + noise * torch.randn((Q, D), generator=g, device="cuda")
    dbo = eng.prepare(db, flags=native.NF_ROWNORM, precision="rescore"); qo = eng.prepare(qs, flags=native.NF_ROWNORM, precision="rescore")
    keys = eng.topk_search(qo, dbo, K, "rescore")
    st = eng.consume_status()
    sc, idx = unpack_keys(keys)
    full = (qo.f32 @ dbo.f32.T)
    top = torch.sort(full, dim=1, descending=True).values[:, :80].cpu().numpy()
    print(n_clusters, per, noise, "status", st, "stats q", qo.stats.tolist(), "db", dbo.stats.tolist())
    print("  mean score rank1 %.4f rank10 %.4f rank16 %.4f rank40 %.4f rank41 %.4f rank64 %.4f" % tuple(top[:, [0, 9, 15, 39, 40, 63]].mean(axis=0)))
