#!/bin/bash
# round 2, visit I (8 GPUs): bench.py --gpus 8 as the driver launches it (multi_parity, weak 4M atoms per GPU, strong_32M leg,
# per-kernel classes), host-side wall trace of the brick step on stderr
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > $O/r02_box_i.txt
PISB_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 40 --warmup 6 > $O/r02_bench_8gpu.log 2> $O/r02_bench_8gpu.err; echo "rc=$?" >> $O/r02_bench_8gpu.log
tail -n 2 $O/r02_bench_8gpu.log | cut -c1-6000
grep "pisb rank 0" $O/r02_bench_8gpu.err | tail -n 12
