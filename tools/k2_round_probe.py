"""Per-round unit durations of the CTA-pair Top-K kernel on the C2 shape (whole database and one 1/8 shard): are the
units of the FIRST round -- which start with empty per-query lists and no shared threshold -- slower than the rest?
Uses the unit stamps of EMR2A_TC_UNIT_CLOCK=1 (emr2a_debug_unit_clocks)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emr2a_b200 import native, synth
from emr2a_b200.engine import get_engine
eng = get_engine(); dev = eng.device; lib = native.load()
flags = native.NF_SEGNORM | native.NF_ROWNORM
n, d, n_q, k = 1_000_000, 512, 10_000, 10
di, _ = synth.device_block(0, n, d, 3, 1234, dev, label_seed=1234); dt, _ = synth.device_block(0, n, d, 3, 1235, dev, label_seed=1234)
qi, _ = synth.device_block(50_003_968, n_q, d, 3, 1234, dev, label_seed=1234); qt, _ = synth.device_block(50_003_968, n_q, d, 3, 1235, dev, label_seed=1234)
qs = eng.prepare(qi, qt, 1.0, 1.0, flags, "rescore")
for parts in (1, 8):
    rows = n // parts
    db = eng.prepare(di[:rows], dt[:rows], 1.0, 1.0, flags, "rescore")
    for _ in range(2):
        eng.topk_filter(qs, db, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.topk_filter(qs, db, k); e1.record(); torch.cuda.synchronize()
    os.environ["EMR2A_TC_UNIT_CLOCK"] = "1"
    eng.topk_filter(qs, db, k); torch.cuda.synchronize()
    os.environ.pop("EMR2A_TC_UNIT_CLOCK")
    cap = 1 << 16
    buf = np.zeros(2 * cap, dtype=np.uint64); plan = np.zeros(8, dtype=np.int64)
    native.check(lib.emr2a_debug_unit_clocks(buf.ctypes.data_as(C.c_void_p), cap, plan.ctypes.data_as(C.c_void_p)))
    m_tiles, n_tiles, splits, tps, mg, st, grid, n_units = [int(x) for x in plan]
    t = buf[:2 * n_units].reshape(n_units, 2).astype(np.float64) / 1e3
    t -= t[:, 0].min()
    dur = t[:, 1] - t[:, 0]
    n_pairs = grid // 2
    print(f"shard 1/{parts}: filter {e0.elapsed_time(e1):.3f} ms; plan m_tiles={m_tiles} n_tiles={n_tiles} splits={splits} tiles/split={tps} units={n_units} pairs={n_pairs}; kernel span {t[:, 1].max():.0f} us")
    for r in range((n_units + n_pairs - 1) // n_pairs):
        sel = np.arange(r * n_pairs, min((r + 1) * n_pairs, n_units))
        print(f"   round {r}: {len(sel):3d} units, duration us min/med/max {dur[sel].min():.0f}/{np.median(dur[sel]):.0f}/{dur[sel].max():.0f}, "
              f"start {t[sel, 0].min():.0f}..{t[sel, 0].max():.0f}, end {t[sel, 1].min():.0f}..{t[sel, 1].max():.0f}")
    del db
