#!/bin/bash
# round 2, call 13 (8 GPUs): the driver's bench command at N=8 (C2 headline + e2e + C5 10M sub-record), stage
# breakdown, bit-identity across shardings, H2D topology probe
O=gpurun_out/r2k
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 870 $TR --master-port 29811 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"; tail -3 $O/bench_n8.err
timeout 300 $TR --master-port 29812 tools/step_breakdown_dist.py > $O/breakdown.log 2>&1; echo "breakdown rc=$?"; grep -v "^\[\|^W\|^\*\|OMP\|^$" $O/breakdown.log | tail -12
timeout 600 $TR --master-port 29813 tools/dist_check.py > $O/dist_check.log 2>&1; echo "dist_check rc=$?"; grep "rank 0" $O/dist_check.log | tail -14
timeout 300 $TR --master-port 29814 tools/h2d_probe.py > $O/h2d_probe.log 2>&1; echo "probe rc=$?"; grep -v "^\[\|^W\|^\*\|OMP\|^$" $O/h2d_probe.log | tail -22
python - <<'PY'
import json
for line in open("gpurun_out/r2k/bench_n8.json"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"e2e",round(d["e2e"]["value"]),d["e2e"].get("ms_per_step"),"unverified",d["unverified_queries"])
        c=d.get("c5") or {}
        print("c5",c.get("value"),c.get("ms_per_step"),(c.get("roofline") or {}).get("frac"),c.get("unverified_queries"),(c.get("cpu_baseline") or {}).get("parity_on_sample"))
PY
