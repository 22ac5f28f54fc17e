#!/bin/bash
# round 2, call 8 (2 GPUs): cooperative shards under NCCL -- bit-identity check, then C2 step with and without them
O=gpurun_out/r2f
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_coop_shards.py -x -q > $O/pytest_coop.log 2>&1; echo "coop rc=$?"; tail -5 $O/pytest_coop.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tools/dist_check.py > $O/dist_check.log 2>&1; echo "dist_check rc=$?"; grep "rank 0" $O/dist_check.log | tail -14
for coop in 1 0; do
  EMR2A_COOP_SHARDS=$coop timeout 900 $TR --master-port 2952$coop bench.py --gpus 2 --no-c5 --no-e2e --steps 50 > $O/bench_n2_coop$coop.json 2> $O/bench_n2_coop$coop.err; echo "bench n2 coop=$coop rc=$?"
done
python - <<'PY'
import json
for c in (1,0):
    for line in open(f"gpurun_out/r2f/bench_n2_coop{c}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print("coop",c,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"unverified",d["unverified_queries"],"launches",d["gpu_launches"])
PY
