#!/bin/bash
O=gpurun_out/r2t
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
i=0
for cfg in "EMR2A_E2E_GATHER_Q=1 EMR2A_HOST_TAPER=0" "EMR2A_E2E_GATHER_Q=0 EMR2A_HOST_TAPER=0" "EMR2A_E2E_GATHER_Q=1 EMR2A_HOST_TAPER=1"; do
  i=$((i+1))
  env $cfg timeout 600 $TR --master-port 2962$i bench.py --gpus 2 --steps 5 --warmup 3 --no-c5 > $O/bench_$i.json 2> $O/bench_$i.err; echo "$cfg rc=$?"
  python - $i <<'PY'
import json,sys
for line in open(f"gpurun_out/r2t/bench_{sys.argv[1]}.json"):
    if line.startswith("{"):
        d=json.loads(line); e=d["e2e"]; print("   e2e",round(e["value"]),round(e["ms_per_step"],2))
PY
done
