#!/bin/bash
# round 2, call 10 (1 GPU): deferred fp32 rows + cooperative shards tests, whole GPU suite, C2 with / without deferred rows, K sweep
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_deferred_rows.py tests/test_gpu_coop_shards.py -x -q > $O/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -15 $O/pytest_new.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
for d in 1 0; do
  EMR2A_DEFER_F32=$d timeout 300 python tools/step_breakdown.py > $O/breakdown_defer$d.log 2>&1; echo "breakdown defer=$d rc=$?"; tail -3 $O/breakdown_defer$d.log
  EMR2A_DEFER_F32=$d timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 30 > $O/bench_c2_defer$d.json 2> $O/bench_c2_defer$d.err; echo "bench defer=$d rc=$?"
done
ARMS="rescore:5,rescore:10,bf16x3:16,bf16x3:32,fp32:64" timeout 600 python tools/k2_sweep.py > $O/k_sweep.log 2>&1; echo "k sweep rc=$?"; cat $O/k_sweep.log | tail -6
python - <<'PY'
import json
for c in (1,0):
    for line in open(f"gpurun_out/r2h/bench_c2_defer{c}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print("defer",c,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"frac",round(r["frac"],3),"unverified",d["unverified_queries"],d["clocks"])
PY
