#!/bin/bash
O=gpurun_out/r2
mkdir -p $O
python tools/k2_cohort_probe.py --clocks > $O/probe_plain.log 2>&1
echo "probe rc=$?"
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:tc2_topk --csv --log-file $O/probe_ncu.csv python tools/k2_cohort_probe.py > $O/probe_ncu.log 2>&1
echo "ncu probe rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?"
tail -5 $O/pytest_gpu.log
