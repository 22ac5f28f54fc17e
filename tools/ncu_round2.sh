#!/bin/bash
# Round-2 ncu captures of the dominant K2 kernel (CTA pairs, query-tile grouping + cohort pacing) on every single-GPU
# workload: C2, C3 (5M rows, late fusion), C4 (D = 5120, bf16 in), the fold-masked C5 instance, the 3-pass arm on C2;
# plus the launch list of one C2 step.  Every ncu run is preceded by the same command exiting 0 without ncu.
set -x
O=gpurun_out/r2b
mkdir -p $O
B="python bench.py --no-e2e --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on -f"
run() {   # name, skip, env..., args...
  local name=$1 skip=$2; shift 2
  env "$@" $B --steps 2 --warmup 3 > $O/${name}_plain.json 2> $O/${name}_plain.err || { echo "plain $name failed"; return; }
  env "$@" $NCU -k regex:tc2_topk -s $skip -c 1 -o $O/prof_${name} $B --steps 1 --warmup 3 > $O/${name}_ncu.log 2>&1
  ncu -i $O/prof_${name}.ncu-rep --page raw --csv > $O/prof_${name}.csv 2>/dev/null
}
run c2 1 EMR2A_BENCH_WORKLOAD=c2
run c4 1 EMR2A_BENCH_WORKLOAD=c4
run c3 1 EMR2A_BENCH_WORKLOAD=c3
run c5 2 EMR2A_BENCH_WORKLOAD=c5 EMR2A_C5_N=2000000
run c2x3 1 EMR2A_BENCH_WORKLOAD=c2 EMR2A_BENCH_PRECISION=bf16x3
# launch list + full capture of every kernel of one C2 step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv $B --steps 2 --warmup 3 > $O/launches_c2.log 2>&1
$NCU -k regex:'normalize_fuse|tc2_topk|topk_merge|rescore_select|exact_rescan|rescan_merge|vote_metrics' -s 32 -c 8 -o $O/prof_step_c2 $B --steps 2 --warmup 3 > $O/step_c2_ncu.log 2>&1
ncu -i $O/prof_step_c2.ncu-rep --page raw --csv > $O/prof_step_c2.csv 2>/dev/null
rm -f $O/prof_step_c2.ncu-rep
ls -la $O
