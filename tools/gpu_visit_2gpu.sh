#!/bin/bash
# round 2, visit K (2 GPUs): multi-GPU equivalence tests (incl. the asynchronous owned download, NVT leg at tau = 100) and
# bench.py --gpus 2 as the driver launches it, on the tree with warp-aggregated sentinel-bucket atomics
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k multi_gpu > $O/r02_pytest_2gpus.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_2gpus.log; tail -6 $O/r02_pytest_2gpus.log
PISB_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 60 --warmup 10 > $O/r02_bench_2gpu.log 2> $O/r02_bench_2gpu.err; echo "rc=$?" >> $O/r02_bench_2gpu.log
tail -n 2 $O/r02_bench_2gpu.log | cut -c1-3500
grep "pisb rank 0" $O/r02_bench_2gpu.err | tail -n 4
