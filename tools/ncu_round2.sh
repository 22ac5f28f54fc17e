#!/bin/bash
# Round-2 ncu captures of the dominant K2 kernel on the workloads the round-1 verdict asked for:
# C4 (D = 5120, bf16 in), C3 (5M rows, late fusion), the fold-masked C5 instance, and the 3-pass arm on C2.
# Every ncu run is preceded by the same command exiting 0 without ncu.  Run under gpurun (one GPU).
set -x
O=gpurun_out/r2
mkdir -p $O
B="python bench.py --no-e2e --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on -f"
run() {   # name, skip, env..., args...
  local name=$1 skip=$2; shift 2
  env "$@" $B --steps 2 --warmup 3 > $O/${name}_plain.json 2> $O/${name}_plain.err || { echo "plain $name failed"; return; }
  env "$@" $NCU -k regex:tc2_topk -s $skip -c 1 -o $O/prof_${name} $B --steps 1 --warmup 3 > $O/${name}_ncu.log 2>&1
  ncu -i $O/prof_${name}.ncu-rep --page raw --csv > $O/prof_${name}.csv 2>/dev/null
}
run c4 1 EMR2A_BENCH_WORKLOAD=c4
run c3 1 EMR2A_BENCH_WORKLOAD=c3
run c5 2 EMR2A_BENCH_WORKLOAD=c5 EMR2A_C5_N=2000000
run c2x3 1 EMR2A_BENCH_WORKLOAD=c2 EMR2A_BENCH_PRECISION=bf16x3
ls -la $O
