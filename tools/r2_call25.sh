#!/bin/bash
O=gpurun_out/r2w
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_deferred_rows.py tests/test_gpu_coop_shards.py tests/test_gpu_kernels.py tests/test_gpu_rescore_bound.py tests/test_gpu_properties.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
timeout 600 python tools/rescore_stage_probe.py > $O/stage_probe.log 2>&1; echo "probe rc=$?"; cat $O/stage_probe.log | tail -4
for dm in 2048 8192; do
  EMR2A_DEFER_MAX_DIM=$dm timeout 600 python bench.py --workload c4 --no-e2e --steps 5 > $O/c4_dm$dm.json 2> $O/c4_dm$dm.err; echo "c4 defer_max=$dm rc=$?"
done
python - <<'PY'
import json
for dm in (2048,8192):
    for line in open(f"gpurun_out/r2w/c4_dm{dm}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print("c4 defer_max",dm,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"frac",round(r["frac"],3),"unverified",d["unverified_queries"],(d["cpu_baseline"] or {}).get("parity_on_sample",{}).get("ok"))
PY
