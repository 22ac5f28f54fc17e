#!/bin/bash
O=gpurun_out/r2d
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
timeout 600 python tools/pp_bench.py > $O/pp_bench.log 2>&1; echo "pp rc=$?"; head -8 $O/pp_bench.log
timeout 600 python bench.py --workload c1 > $O/bench_c1.json 2> $O/bench_c1.err; echo "c1 rc=$?"
