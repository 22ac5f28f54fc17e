#!/bin/bash
O=gpurun_out/r2
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu2.log
for w in c2 c2k5 c3 c4; do
  timeout 900 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"
done
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload c1 > $O/bench_c1.json 2> $O/bench_c1.err; echo "c1 rc=$?"
