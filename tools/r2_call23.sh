#!/bin/bash
# round 2, call 23 (8 GPUs): C2 headline + e2e at HEAD (query gather over NVLink, H2D-weighted shards), no C5 record
O=gpurun_out/r2u
mkdir -p $O
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR8 --master-port 29921 bench.py --gpus 8 --steps 20 --warmup 5 --no-c5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"; tail -2 $O/bench_n8.err
python - <<'PY'
import json
for line in open("gpurun_out/r2u/bench_n8.json"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]; e=d["e2e"]
        print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"traffic",r["traffic"],"e2e",round(e["value"]),round(e["ms_per_step"],2),e.get("h2d_bytes_per_step"),e.get("shards"),"unverified",d["unverified_queries"])
PY
