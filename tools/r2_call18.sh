#!/bin/bash
# round 2, call 18 (8-GPU box): the driver's bench command at N=8 (weighted host shards for e2e), at N=4 on the SAME box
# (ranks spread over devices 0,2,4,6), stage breakdown at N=8
O=gpurun_out/r2p
mkdir -p $O
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 870 $TR8 --master-port 29911 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"; tail -2 $O/bench_n8.err
timeout 600 $TR4 --master-port 29912 bench.py --gpus 4 --steps 20 --warmup 5 --no-c5 > $O/bench_n4_on8.json 2> $O/bench_n4_on8.err; echo "bench n4 rc=$?"; tail -2 $O/bench_n4_on8.err
timeout 300 $TR8 --master-port 29913 tools/step_breakdown_dist.py > $O/breakdown.log 2>&1; echo "breakdown rc=$?"; grep -v "^\[\|^W\|^\*\|OMP\|^$" $O/breakdown.log | tail -11
python - <<'PY'
import json
for f in ("bench_n8","bench_n4_on8"):
    for line in open(f"gpurun_out/r2p/{f}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]; e=d["e2e"]
            print(f,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"e2e",round(e["value"]),round(e["ms_per_step"],2),e.get("shards"),"unverified",d["unverified_queries"])
            c=d.get("c5") or {}
            if c: print("  c5",c.get("value"),c.get("ms_per_step"),(c.get("roofline") or {}).get("frac"),c.get("unverified_queries"),((c.get("cpu_baseline") or {}).get("parity_on_sample") or {}).get("ok"))
PY
