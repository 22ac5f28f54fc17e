#!/bin/bash
# round 2, call 7 (1 GPU): cooperative-shard stage kernels (emulated shards) + whole GPU suite + C2 bench line
O=gpurun_out/r2e
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_coop_shards.py -x -q > $O/pytest_coop.log 2>&1; echo "coop rc=$?"; tail -15 $O/pytest_coop.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
