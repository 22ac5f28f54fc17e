"""K2 (CTA-pair kernel) database-stream sharing probe: unit grouping (EMR2A_TC_A_MB) and cohort pacing
(EMR2A_TC_SYNC / EMR2A_TC_SYNC_BUDGET) on C3-, C4- and C5-shaped problems.

    python tools/k2_cohort_probe.py [--shapes c3,c4,c5] [--clocks]

Prints one line per (shape, setting): K2 milliseconds (CUDA events, best of reps) and TFLOP/s.  Run the same command
under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:tc2_topk` for the DRAM bytes per launch
(launch order = print order x reps).  --clocks adds per-unit start/end stamps (emr2a_debug_unit_clocks): spread of the
start times inside a cohort and of the unit durations."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SETTINGS = [
    ("r1 order, no pacing", {"EMR2A_TC_A_MB": "100000", "EMR2A_TC_SYNC": "0"}),
    ("grouped, no pacing", {"EMR2A_TC_SYNC": "0"}),
    ("grouped + pacing (default)", {}),
    ("grouped + pacing S=2", {"EMR2A_TC_SYNC": "2"}),
    ("grouped + pacing S=8", {"EMR2A_TC_SYNC": "8"}),
    ("grouped + pacing budget 40k", {"EMR2A_TC_SYNC_BUDGET": "40000"}),
    ("A budget 12 MB + pacing", {"EMR2A_TC_A_MB": "12"}),
    ("A budget 40 MB + pacing", {"EMR2A_TC_A_MB": "40"}),
]
KNOBS = ("EMR2A_TC_A_MB", "EMR2A_TC_SYNC", "EMR2A_TC_SYNC_BUDGET", "EMR2A_TC_UNIT_CLOCK")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="c3,c4,c5")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--clocks", action="store_true")
    ap.add_argument("--settings", default="")
    args = ap.parse_args()
    import torch
    from emr2a_b200 import native, synth
    from emr2a_b200.engine import get_engine
    eng = get_engine()
    dev = eng.device
    lib = native.load()
    pick = [int(x) for x in args.settings.split(",")] if args.settings else range(len(SETTINGS))

    def clocks_report(tag):
        cap = 1 << 16
        buf = np.zeros(2 * cap, dtype=np.uint64)
        plan = np.zeros(8, dtype=np.int64)
        native.check(lib.emr2a_debug_unit_clocks(buf.ctypes.data_as(C.c_void_p), cap, plan.ctypes.data_as(C.c_void_p)))
        m_tiles, n_tiles, splits, tps, mg, st, grid, n_units = [int(x) for x in plan]
        n_units = min(n_units, cap)
        t = buf[:2 * n_units].reshape(n_units, 2).astype(np.float64) / 1e3      # us
        t -= t[:, 0].min()
        dur = t[:, 1] - t[:, 0]
        n_pairs = grid // 2
        per_group = mg * splits
        spreads = []
        for u0 in range(0, n_units):
            g = u0 // per_group
            r = u0 - g * per_group
            mg_eff = min(mg, m_tiles - g * mg)
            if r % mg_eff:
                continue
            members = np.arange(u0, min(u0 + mg_eff, n_units))
            for step in np.unique(members // n_pairs):
                part = members[members // n_pairs == step]
                if len(part) > 1:
                    spreads.append((t[part, 0].max() - t[part, 0].min(), t[part, 1].max() - t[part, 1].min()))
        sp = np.array(spreads) if spreads else np.zeros((1, 2))
        print(f"    [{tag}] plan m_tiles={m_tiles} n_tiles={n_tiles} splits={splits} tiles/split={tps} mg={mg} sync={st} units={n_units}; "
              f"unit duration us min/med/max {dur.min():.0f}/{np.median(dur):.0f}/{dur.max():.0f}; cohort start spread us med/max "
              f"{np.median(sp[:, 0]):.0f}/{sp[:, 0].max():.0f}; cohort end spread us med/max {np.median(sp[:, 1]):.0f}/{sp[:, 1].max():.0f}; "
              f"last end {t[:, 1].max():.0f} us", flush=True)

    def run(name, q, db, k, prec, flops, **kw):
        for i in pick:
            label, env = SETTINGS[i]
            for kn in KNOBS:
                os.environ.pop(kn, None)
            os.environ.update(env)
            best = 1e9
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.topk_search(q, db, k, prec, **kw)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            eng.consume_status()
            print(f"{name:4s} {label:32s} {best:9.3f} ms  {flops / best / 1e9:8.1f} TFLOP/s", flush=True)
            if args.clocks:
                os.environ["EMR2A_TC_UNIT_CLOCK"] = "1"
                eng.topk_search(q, db, k, prec, **kw)
                torch.cuda.synchronize()
                eng.consume_status()
                clocks_report(label)
        for kn in KNOBS:
            os.environ.pop(kn, None)

    shapes = args.shapes.split(",")
    flags = native.NF_SEGNORM | native.NF_ROWNORM
    if "c3" in shapes:
        n, d, n_q, k = 5_000_000, 512, 10_000, 10
        di, _ = synth.device_block(0, n, d, 3, 13, dev, label_seed=13)
        dt, _ = synth.device_block(0, n, d, 3, 14, dev, label_seed=13)
        qi, _ = synth.device_block(50_003_968, n_q, d, 3, 13, dev, label_seed=13)
        qt, _ = synth.device_block(50_003_968, n_q, d, 3, 14, dev, label_seed=13)
        db = eng.prepare(di, dt, 1.0, 1.0, native.NF_SEGNORM, "bf16x1")
        q = eng.prepare(qi, qt, 0.75, 0.25, native.NF_SEGNORM, "bf16x1")
        del di, dt
        run("c3", q, db, k, "bf16x1", 2.0 * 1024 * n_q * n)
        del db, q
        torch.cuda.empty_cache()
    if "c4" in shapes:
        n, n_q, k = 1_000_000, 10_000, 10
        bf = torch.bfloat16
        di, _ = synth.device_block(0, n, 4096, 3, 17, dev, label_seed=17, dtype=bf)
        dt, _ = synth.device_block(0, n, 1024, 3, 18, dev, label_seed=17, dtype=bf)
        qi, _ = synth.device_block(50_003_968, n_q, 4096, 3, 17, dev, label_seed=17, dtype=bf)
        qt, _ = synth.device_block(50_003_968, n_q, 1024, 3, 18, dev, label_seed=17, dtype=bf)
        db = eng.prepare(di, dt, 1.0, 1.0, flags, "bf16x1")
        q = eng.prepare(qi, qt, 1.0, 1.0, flags, "bf16x1")
        del di, dt
        run("c4", q, db, k, "bf16x1", 2.0 * 5120 * n_q * n)
        del db, q
        torch.cuda.empty_cache()
    if "c5" in shapes:
        n, d, k, n_folds, qb = 2_000_000, 512, 5, 5, 131072
        di, _ = synth.device_block(0, n, d, 3, 19, dev, label_seed=19)
        dt, _ = synth.device_block(0, n, d, 3, 20, dev, label_seed=19)
        fold = (torch.arange(n, device=dev, dtype=torch.int64) * n_folds // n).to(torch.uint8)
        db = eng.prepare(di, dt, 1.0, 1.0, flags, "bf16x1")
        del di, dt
        q = eng._rows(db, 0, qb)                       # queries = the first block of the cohort (all in fold 0)
        run("c5", q, db, k, "bf16x1", 2.0 * 1024 * qb * (n - n // n_folds), q_fold=fold[:qb], db_fold=fold, fold_sorted=True)


if __name__ == "__main__":
    main()
