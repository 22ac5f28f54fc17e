#!/bin/bash
# round 2, call 20 (1 GPU): what the driver runs at round end (smoke, GPU tests, bench both arms) + the captures of the
# final kernels (launch list and ncu --set full of every kernel of a C2 step) + serving latencies
O=gpurun_out/r2r
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout 900 python bench.py --workload c2k5 --no-e2e > $O/bench_c2k5.json 2> $O/bench_c2k5.err; echo "c2k5 rc=$?"
B="python bench.py --no-e2e --no-cpu-baseline"
$B --steps 2 --warmup 3 > $O/c2_plain.json 2> $O/c2_plain.err && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv $B --steps 2 --warmup 3 > $O/launches_c2.log 2>&1
  ncu --set full --clock-control none --import-source on -f -k regex:'normalize_fuse|tc2_topk|topk_merge|rescore_select|exact_rescan|rescan_merge|vote_metrics' -s 32 -c 8 -o $O/prof_step_c2 $B --steps 2 --warmup 3 > $O/step_c2_ncu.log 2>&1
  ncu -i $O/prof_step_c2.ncu-rep --page raw --csv > $O/prof_step_c2.csv 2>/dev/null
  rm -f $O/prof_step_c2.ncu-rep
}
PRECS=rescore QS=1,64,256,1024 timeout 300 python tools/latency_bench.py > $O/latency.log 2>&1; echo "latency rc=$?"; tail -4 $O/latency.log
python - <<'PY'
import json
for w in ("c2","c2k5","ref"):
    for line in open(f"gpurun_out/r2r/bench_{w}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d.get("roofline") or {}
            print(w,"value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"k2_ms",r.get("kernel_ms"),"frac",r.get("frac"),"traffic",r.get("traffic"),"e2e",(d.get("e2e") or {}).get("value"),"unverified",d.get("unverified_queries"),"launches",d.get("gpu_launches"))
PY
