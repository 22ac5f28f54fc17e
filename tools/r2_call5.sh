#!/bin/bash
O=gpurun_out/r2c
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
(numactl -H; lscpu | head -30; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class; then echo "$d $(cat $d/local_cpulist) numa $(cat $d/numa_node)"; fi; done) > $O/numa.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > $O/dist_check.log 2>&1; echo "dist_check rc=$?"; grep "rank 0" $O/dist_check.log | tail -12
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
tail -3 $O/bench_n2.err
