#!/bin/bash
# usage: gpu_multi.sh NPROC [ncell steps T0]
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_check.py ${2:-16} ${3:-60} ${4:-60} > gpurun_out/multi_check_$N.log 2>&1
echo "rc=$?" >> gpurun_out/multi_check_$N.log
tail -n 25 gpurun_out/multi_check_$N.log | cut -c1-1500
