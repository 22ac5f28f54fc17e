#!/bin/bash
# round 2, final check at HEAD (1 GPU): smoke, whole GPU suite, both bench arms, ncu of the filtered re-scan on C4
O=gpurun_out/r2z
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:'filtered_rescan|rescore_select' -s 6 -c 2 -o $O/prof_c4_rescan python bench.py --workload c4 --no-e2e --no-cpu-baseline --steps 1 --warmup 3 > $O/c4_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_c4_rescan.ncu-rep --page raw --csv > $O/prof_c4_rescan.csv 2>/dev/null; rm -f $O/prof_c4_rescan.ncu-rep
python - <<'PY'
import json
for w in ("c2","ref"):
    for line in open(f"gpurun_out/r2z/bench_{w}.json"):
        if line.startswith("{"):
            d=json.loads(line); r=d.get("roofline") or {}
            print(w,"value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"k2_ms",r.get("kernel_ms"),"frac",r.get("frac"),"traffic",r.get("traffic"),"e2e",(d.get("e2e") or {}).get("value"),"unverified",d.get("unverified_queries"),"parity",((d.get("cpu_baseline") or {}).get("parity_on_sample") or {}).get("ok"))
PY
