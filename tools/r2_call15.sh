#!/bin/bash
O=gpurun_out/r2m
mkdir -p $O
timeout 600 python tools/rescore_stage_probe.py > $O/stage_probe.log 2>&1; echo "probe rc=$?"; cat $O/stage_probe.log | tail -6
timeout 600 python -m pytest tests/test_gpu_deferred_rows.py tests/test_gpu_coop_shards.py tests/test_gpu_kernels.py -x -q -k "rescore or deferred or coop or flagged" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
