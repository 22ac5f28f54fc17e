#!/bin/bash
# round 2, call 21 (2 GPUs): host-path tests, then the driver's bench command at N=2 (query gather + tapered chunks in e2e)
O=gpurun_out/r2s
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_host_path.py tests/test_gpu_kernels.py -q -k "host_path or duplicate_rows" > $O/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -4 $O/pytest_new.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 870 $TR --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -2 $O/bench_n2.err
python - <<'PY'
import json
for f in ("bench_n1","bench_n2"):
    for line in open(f"gpurun_out/r2s/{f}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]; e=d["e2e"]
            print(f,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"k2_ms",round(r["kernel_ms"],3),"e2e",round(e["value"]),round(e["ms_per_step"],2),e.get("h2d_bytes_per_step"),e.get("shards"),"unverified",d["unverified_queries"])
            c=d.get("c5") or {}
            if c: print("  c5",c.get("value"),c.get("ms_per_step"),(c.get("roofline") or {}).get("frac"),c.get("unverified_queries"),((c.get("cpu_baseline") or {}).get("parity_on_sample") or {}).get("ok"))
PY
