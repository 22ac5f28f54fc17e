#!/bin/bash
# round 2, call 9 (8 GPUs): cooperative shards at N=8 -- stage breakdown, A/B of the C2 step, bit-identity
O=gpurun_out/r2g
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for coop in 1 0; do
  EMR2A_COOP_SHARDS=$coop timeout 300 $TR --master-port 2961$coop tools/step_breakdown_dist.py > $O/breakdown_coop$coop.log 2>&1; echo "breakdown coop=$coop rc=$?"; grep -v "^\[\|^W\|^\*" $O/breakdown_coop$coop.log | tail -14
done
for coop in 1 0; do
  EMR2A_COOP_SHARDS=$coop timeout 600 $TR --master-port 2962$coop bench.py --gpus 8 --no-c5 --no-e2e --steps 100 > $O/bench_n8_coop$coop.json 2> $O/bench_n8_coop$coop.err; echo "bench n8 coop=$coop rc=$?"
done
python - <<'PY'
import json
for c in (1,0):
    for line in open(f"gpurun_out/r2g/bench_n8_coop{c}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print("coop",c,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"unverified",d["unverified_queries"],"launches",d["gpu_launches"])
PY
timeout 600 $TR --master-port 29631 tools/dist_check.py > $O/dist_check.log 2>&1; echo "dist_check rc=$?"; grep "rank 0" $O/dist_check.log | tail -14
