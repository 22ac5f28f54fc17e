#!/bin/bash
# round 2, call 11 (1 GPU): progressive cut + deferred rows -- tests, per-stage times, C2 line with / without deferred rows
O=gpurun_out/r2i
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_deferred_rows.py tests/test_gpu_coop_shards.py tests/test_gpu_rescore_bound.py -x -q > $O/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -5 $O/pytest_new.log
for d in 1 0; do
  EMR2A_DEFER_F32=$d timeout 300 python tools/step_breakdown.py > $O/breakdown_defer$d.log 2>&1; echo "breakdown defer=$d rc=$?"; tail -2 $O/breakdown_defer$d.log
  EMR2A_DEFER_F32=$d timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 30 > $O/bench_c2_defer$d.json 2> $O/bench_c2_defer$d.err; echo "bench defer=$d rc=$?"
done
for d in 1 0; do
  EMR2A_DEFER_F32=$d ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'rescore_select|normalize_fuse' -c 12 --csv --log-file $O/ncu_small_defer$d.csv python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 3 > $O/ncu_small_defer$d.log 2>&1; echo "ncu defer=$d rc=$?"
done
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
python - <<'PY'
import json, csv
for c in (1,0):
    for line in open(f"gpurun_out/r2i/bench_c2_defer{c}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print("defer",c,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"frac",round(r["frac"],3),"unverified",d["unverified_queries"],d["clocks"]["sm_mhz"])
    rows=[r for r in csv.reader(open(f"gpurun_out/r2i/ncu_small_defer{c}.csv", errors="replace")) if len(r)>10]
    h=rows[0]; 
    for r in rows[-9:]:
        print("  ", r[h.index("Kernel Name")][:60], r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
