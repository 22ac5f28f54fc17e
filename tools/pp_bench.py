"""Per-fold preprocessing and fused late fusion (8f-4) alone: achieved HBM GB/s of the two
hand-written passes, fit/transform wall time, and the z-score / min-max late search next to the plain one."""
import json, os, sys, time, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from emr2a_b200 import native, synth, preprocess as pp
from emr2a_b200.engine import get_engine
from emr2a_b200.late import late_fusion_search
eng = get_engine(); dev = eng.device
PEAK = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6549.4
n, d, p = int(os.environ.get("N", 2_000_000)), int(os.environ.get("D", 512)), int(os.environ.get("P", 128))
x, _ = synth.device_block(0, n, d, 3, 11, dev)
x = x * 3.0 + 0.7


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: pp.column_moments(eng, x))
gb = n * d * 4 / 1e9
print(f"column_moments {n}x{d}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s ({gb:.2f} GB read) = {gb/ms*1e3/PEAK:.2f} of measured HBM peak", flush=True)
tf = pp.fit_scaler(eng, x)
out = torch.empty_like(x)
ms = timed(lambda: pp.standardize(eng, x, tf.mean_f32, tf.scale_f32, out=out))
gb = 2 * n * d * 4 / 1e9
print(f"standardize {n}x{d}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s ({gb:.2f} GB read+write) = {gb/ms*1e3/PEAK:.2f} of measured HBM peak", flush=True)
del out
if os.environ.get("KERNELS_ONLY") == "1":
    sys.exit(0)
torch.cuda.synchronize(); t0 = time.perf_counter()
tf = pp.fit(x, p, eng); torch.cuda.synchronize(); t_fit0 = time.perf_counter() - t0
t0 = time.perf_counter(); tf = pp.fit(x, p, eng); torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
t0 = time.perf_counter(); y = pp.transform(tf, x, eng); torch.cuda.synchronize(); t_tr = time.perf_counter() - t0
print(f"fit (scaler + exact PCA {d}->{p}) on {n} rows: {t_fit*1e3:.1f} ms (first call {t_fit0*1e3:.0f} ms); transform + row-normalise: {t_tr*1e3:.1f} ms", flush=True)
del x, y
# ---- late fusion, 1M x (512 + 512), 10k queries, K = 10
n_db, n_q, dm, k = int(os.environ.get("NDB", 1_000_000)), 10_000, 512, 10
dt, _ = synth.device_block(0, n_db + n_q, dm, 3, 21, dev)
di, _ = synth.device_block(0, n_db + n_q, dm, 3, 22, dev)
args = (dt[:n_db], di[:n_db], dt[n_db:], di[n_db:])
for name, mode in (("none", native.SCORE_NONE), ("zscore", native.SCORE_ZSCORE), ("minmax", native.SCORE_MINMAX)):
    ms = timed(lambda: late_fusion_search(*args, 0.4, mode, k, engine=eng), reps=3, warm=1)
    print(f"late fusion '{name}' {n_db}x({dm}+{dm}), {n_q} queries, K={k}: {ms:.1f} ms/step = {n_q/ms*1e3:.0f} queries/s", flush=True)
from emr2a_b200.late import LateFusionIndex
index = LateFusionIndex(args[0], args[1], k=k, expected_queries=n_q, engine=eng)
for name, mode in (("none", native.SCORE_NONE), ("zscore", native.SCORE_ZSCORE), ("minmax", native.SCORE_MINMAX)):
    ms = timed(lambda: index.search(args[2], args[3], 0.4, mode, k), reps=3, warm=1)
    print(f"resident LateFusionIndex '{name}': {ms:.1f} ms/step = {n_q/ms*1e3:.0f} queries/s", flush=True)
