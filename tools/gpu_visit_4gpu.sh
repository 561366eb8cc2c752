#!/bin/bash
# 4-GPU visit (2x2x1 bricks): bench.py --gpus 4 as the driver launches it
mkdir -p gpurun_out
O=gpurun_out
PISB_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 40 --warmup 6 > $O/r02_bench_4gpu.log 2> $O/r02_bench_4gpu.err; echo "rc=$?" >> $O/r02_bench_4gpu.log
tail -n 2 $O/r02_bench_4gpu.log | cut -c1-3000
