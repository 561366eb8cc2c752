#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
./tools/gpu_multi.sh $N 16 60 60
PISB_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps ${2:-40} --warmup 10 > gpurun_out/bench_multi_$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_multi_$N.log
grep -E "pisb rank|rc=" gpurun_out/bench_multi_$N.log | head -4; tail -n 2 gpurun_out/bench_multi_$N.log | cut -c1-3500
