#!/bin/bash
# round 2, visit F (2 GPUs): the multi-GPU equivalence tests (both halo paths, both step kernels, incl. the device-side
# velocity initialisation on bricks) and bench.py --gpus 2 as the driver launches it (multi_parity + weak leg + strong_32M leg)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader > $O/r02_box_f.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k multi_gpu > $O/r02_pytest_2gpus.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_2gpus.log; tail -6 $O/r02_pytest_2gpus.log
PISB_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 60 --warmup 10 > $O/r02_bench_2gpu.log 2> $O/r02_bench_2gpu.err; echo "rc=$?" >> $O/r02_bench_2gpu.log
tail -n 2 $O/r02_bench_2gpu.log | cut -c1-3000
tail -n 12 $O/r02_bench_2gpu.err
