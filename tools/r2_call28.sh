#!/bin/bash
O=gpurun_out/r3b
mkdir -p $O
timeout 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python - <<'PY'
import json
for line in open("gpurun_out/r3b/bench_c2.json"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("c2 value",round(d["value"]),"ms",round(d["ms_per_step"],3),"kernel_ms",round(r["kernel_ms"],3),"frac",round(r["frac"],3),"search_ms",round(r["search_ms"],3),"e2e",round(d["e2e"]["value"]),"parity",d["cpu_baseline"]["parity_on_sample"]["ok"],"cpu",round(d["cpu_baseline"]["value"],1))
PY
