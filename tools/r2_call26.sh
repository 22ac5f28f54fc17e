#!/bin/bash
O=gpurun_out/r2x
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_deferred_rows.py tests/test_gpu_coop_shards.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_host_path.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 600 python bench.py --workload c4 --steps 5 > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
for line in open("gpurun_out/r2x/bench_c4.json"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("c4 value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"frac",round(r["frac"],3),"e2e",round(d["e2e"]["value"]),"unverified",d["unverified_queries"],(d["cpu_baseline"] or {}).get("parity_on_sample",{}).get("ok"))
PY
