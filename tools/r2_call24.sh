#!/bin/bash
O=gpurun_out/r2v
mkdir -p $O
for kp in 16 32; do
  EMR2A_RESCORE_KP=$kp timeout 600 python bench.py --workload c4 --no-e2e --no-cpu-baseline --steps 5 > $O/c4_kp$kp.json 2> $O/c4_kp$kp.err; echo "c4 kp=$kp rc=$?"
  EMR2A_RESCORE_KP=$kp timeout 600 python bench.py --workload c2 --no-e2e --no-cpu-baseline --steps 20 > $O/c2_kp$kp.json 2> $O/c2_kp$kp.err; echo "c2 kp=$kp rc=$?"
done
python - <<'PY'
import json
for w in ("c4","c2"):
    for kp in (16,32):
        for line in open(f"gpurun_out/r2v/{w}_kp{kp}.json"):
            if line.startswith("{"):
                d=json.loads(line); r=d["roofline"]
                print(w,"kp",kp,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"unverified",d["unverified_queries"],"steps",d["steps"])
PY
