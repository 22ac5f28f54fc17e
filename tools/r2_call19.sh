#!/bin/bash
O=gpurun_out/r2q
mkdir -p $O
python tools/shard_rescore_probe.py > $O/probe.log 2>&1; echo "probe rc=$?"; tail -3 $O/probe.log
ncu --set full --clock-control none --import-source on -f -k regex:rescore_select -s 3 -c 1 -o $O/prof_f32 python tools/shard_rescore_probe.py > $O/ncu_f32.log 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:rescore_select -s 27 -c 1 -o $O/prof_lazy python tools/shard_rescore_probe.py > $O/ncu_lazy.log 2>&1
for v in f32 lazy; do
  ncu -i $O/prof_$v.ncu-rep --page raw --csv > $O/prof_$v.csv 2>/dev/null
  ncu -i $O/prof_$v.ncu-rep --page source --csv > $O/src_$v.csv 2>/dev/null
done
python - <<'PY'
import csv
for v in ("f32","lazy"):
    rows=list(csv.reader(open(f"gpurun_out/r2q/prof_{v}.csv", errors="replace")))
    h=rows[0]
    want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","smsp__inst_executed.sum","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","launch__grid_size","launch__waves_per_multiprocessor","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct"]
    for r in rows[2:]:
        print("----",v)
        for w in want:
            if w in h: print("  ",w[:75], r[h.index(w)][:60])
PY
