#!/bin/bash
mkdir -p gpurun_out
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_check.py 24 60 60 > gpurun_out/multi_check_$N.log 2>&1
echo "rc=$?" >> gpurun_out/multi_check_$N.log
tail -n 3 gpurun_out/multi_check_$N.log | cut -c1-1200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_multi_$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_multi_$N.log
tail -n 2 gpurun_out/bench_multi_$N.log | cut -c1-3000
