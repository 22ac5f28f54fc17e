#!/bin/bash
# round 2, call 12 (4 GPUs): H2D topology probe + the driver's bench command at N=4
O=gpurun_out/r2j
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 $TR --master-port 29711 tools/h2d_probe.py > $O/h2d_probe.log 2>&1; echo "probe rc=$?"; grep -v "^\[\|^W\|^\*\|OMP" $O/h2d_probe.log | tail -14
timeout 870 $TR --master-port 29712 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?"; tail -3 $O/bench_n4.err
python - <<'PY'
import json
for line in open("gpurun_out/r2j/bench_n4.json"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"e2e",round(d["e2e"]["value"]),d["e2e"].get("ms_per_step"),"unverified",d["unverified_queries"])
        c=d.get("c5") or {}
        print("c5",c.get("value"),c.get("ms_per_step"),(c.get("roofline") or {}).get("frac"),c.get("unverified_queries"),(c.get("cpu_baseline") or {}).get("parity"))
PY
