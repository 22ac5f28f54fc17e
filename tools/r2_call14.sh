#!/bin/bash
# round 2, call 14 (1 GPU): final kernels -- tests, bench lines of C2/C3/C4 (deferred rows, progressive cut), ncu captures:
# every kernel of a C2 step, launch list, and the Top-K kernel a rank of a 2/4/8-GPU run launches (--shard-of)
O=gpurun_out/r2l
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_deferred_rows.py tests/test_gpu_coop_shards.py -x -q > $O/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -3 $O/pytest_new.log
timeout 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
B="python bench.py --no-e2e --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on -f"
$B --steps 2 --warmup 3 > $O/c2_plain.json 2> $O/c2_plain.err && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv $B --steps 2 --warmup 3 > $O/launches_c2.log 2>&1
  $NCU -k regex:'normalize_fuse|tc2_topk|topk_merge|rescore_select|exact_rescan|rescan_merge|vote_metrics' -s 32 -c 8 -o $O/prof_step_c2 $B --steps 2 --warmup 3 > $O/step_c2_ncu.log 2>&1
  ncu -i $O/prof_step_c2.ncu-rep --page raw --csv > $O/prof_step_c2.csv 2>/dev/null
}
for n in 2 4 8; do
  $B --steps 2 --warmup 3 --shard-of $n > $O/shard${n}_plain.json 2> $O/shard${n}_plain.err && {
    $NCU -k regex:tc2_topk -s 1 -c 1 -o $O/prof_shard$n $B --steps 1 --warmup 3 --shard-of $n > $O/shard${n}_ncu.log 2>&1
    ncu -i $O/prof_shard$n.ncu-rep --page raw --csv > $O/prof_shard$n.csv 2>/dev/null
    rm -f $O/prof_shard$n.ncu-rep
  }
done
for w in c3 c4; do timeout 900 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; done
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
python - <<'PY'
import json
for w in ("c2","c3","c4"):
    for line in open(f"gpurun_out/r2l/bench_{w}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print(w,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"frac",round(r["frac"],3),r["peak_source"][-40:],"e2e",round(d["e2e"]["value"]),"unverified",d["unverified_queries"],"cpu",round(d["cpu_baseline"]["value"],2),d["cpu_baseline"]["parity_on_sample"]["ok"])
PY
