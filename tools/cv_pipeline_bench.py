"""Whole reference pipeline at scale on one GPU: CVRetrievalEvaluator.run_cv_arrays = StratifiedKFold (host) +
per-fold StandardScaler + exact PCA + fusion + Top-K + votes (device).  N x (512 + 512) -> pca_dim 128 each."""
import os, sys, time, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import logging
from emr2a_b200 import synth
from emr2a_b200.engine import get_engine
from emr2a_b200.utils.cv_evaluator import CVRetrievalEvaluator
logging.getLogger().setLevel(logging.WARNING)
eng = get_engine(); dev = eng.device
n, d, p, k = int(os.environ.get("N", 1_000_000)), int(os.environ.get("D", 512)), int(os.environ.get("P", 128)), 5
img, lab = synth.device_block(0, n, d, 3, 11, dev, label_seed=11)
txt, _ = synth.device_block(0, n, d, 3, 12, dev, label_seed=11)
labels = lab.cpu().numpy()
ev = CVRetrievalEvaluator(cv_folds=5, pca_dim=p, top_k=k, seed=42)
for it in range(2):
    torch.cuda.synchronize(); l0 = eng.launches; t0 = time.perf_counter()
    res = ev.run_cv_arrays(labels, img, txt, fusion=os.environ.get("FUSION", "concat"), top_k_list=[1, 3, 5, 5])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    s = res["summary"]
    print(f"run {it}: full 5-fold CV pipeline (split + scaler + PCA {d}->{p} per modality + search + vote) over {n} cases: "
          f"{dt:.2f} s = {n/dt:.0f} queries/s; top1 {s['top1']['mean']:.4f} vote {s['vote_acc']['mean']:.4f}; "
          f"launches {eng.launches - l0}", flush=True)
