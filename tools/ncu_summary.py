"""Turn ncu CSV exports into the markdown summaries kept under profiles/.

    launch list:  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>
                  python tools/ncu_summary.py launches launches.csv > profiles/rNN_launches_*.md
    full capture: ncu -i capture.ncu-rep --page raw --csv > raw.csv
                  python tools/ncu_summary.py full raw.csv > profiles/rNN_ncu_*.md
"""
import csv
import sys

OURS = ("normalize_fuse", "tc2_topk", "tc_topk", "simt_topk", "topk_merge", "rescore_select", "exact_rescan", "rescan_merge",
        "vote_metrics", "column_moments", "standardize", "scale_segments", "keys_add_offset", "zero_keys", "segment_mean")


def short(name: str) -> str:
    return name.split("(")[0].replace("emr2a::", "")


def launches(path: str) -> None:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 6]
    head = rows[0]
    i_name, i_val = head.index("Kernel Name"), head.index("Metric Value")
    recs = [(short(r[i_name]), float(r[i_val].replace(",", ""))) for r in rows[1:] if r[i_val].replace(",", "").replace(".", "").isdigit()]
    recs = [r for r in recs if any(o in r[0] for o in OURS)]
    # last step = launches after the last database-sized normalize_fuse
    big = max(v for n, v in recs if "normalize_fuse" in n)
    starts = [i for i, (n, v) in enumerate(recs) if "normalize_fuse" in n and v > 0.5 * big]
    step = recs[starts[-1]:]
    total = sum(v for _, v in step)
    print("| # | kernel | ns | share of step |\n|---|---|---|---|")
    for i, (n, v) in enumerate(step):
        print(f"| {i} | `{n}` | {v:.0f} | {100 * v / total:.2f}% |")
    search = sum(v for n, v in step if not ("normalize_fuse" in n or "vote_metrics" in n))
    print(f"\nSum {total / 1e6:.3f} ms; `emr2a_topk_search` (filter + merge + rescore + re-scan launches) = {100 * search / total:.1f}% of the step; "
          f"{len(recs)} launches of ours in the whole run.")


def full(path: str) -> None:
    rows = list(csv.reader(open(path, errors="replace")))
    head, units = rows[0], rows[1]
    want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"),
            ("dram__bytes_write.sum", "DRAM write"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
            ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
            ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
            ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
            ("l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "L2->SM % of peak"),
            ("sm__cycles_elapsed.avg.per_second", "SM clock"), ("launch__registers_per_thread", "regs/thread"),
            ("launch__grid_size", "grid"), ("launch__block_size", "block"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %")]
    cols = [(head.index(k), label) for k, label in want if k in head]
    print("| " + " | ".join(label for _, label in cols) + " |")
    print("|" + "---|" * len(cols))
    for r in rows[2:]:
        if not any(o in r[cols[0][0]] for o in OURS):
            continue
        cells = [f"`{short(r[i])}`" if label == "kernel" else f"{r[i]} {units[i]}".strip() for i, label in cols]
        print("| " + " | ".join(cells) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
