#!/bin/bash
O=gpurun_out/r2n
mkdir -p $O
ncu --set full --clock-control none --import-source on -f -k regex:exact_rescan -c 12 -o $O/prof_rescan python tools/rescore_stage_probe.py > $O/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_rescan.ncu-rep --page raw --csv > $O/prof_rescan.csv 2>/dev/null
ncu -i $O/prof_rescan.ncu-rep --page details --csv > $O/prof_rescan_details.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2n/prof_rescan.csv", errors="replace")))
h=rows[0]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__inst_executed.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio"]
idx=[(h.index(w),w) for w in want if w in h]
for r in rows[2:]:
    print("----")
    for i,w in idx: print("  ",w[:70], r[i][:70])
PY
