#!/bin/bash
# 8-GPU visit (BASELINE configs[3]: 32M atoms, 2x2x2 bricks): the weak-scaling bench leg only.
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 60 --warmup 6 --e2e-steps 20 > gpurun_out/bench_multi_8.log 2>&1; echo "rc=$?" >> gpurun_out/bench_multi_8.log
tail -n 2 gpurun_out/bench_multi_8.log | cut -c1-3000
