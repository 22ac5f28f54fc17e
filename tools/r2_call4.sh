#!/bin/bash
O=gpurun_out/r2b
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
bash tools/ncu_round2.sh > $O/ncu.log 2>&1; echo "ncu rc=$?"
