#!/bin/bash
# 2-GPU visit: the multi-GPU equivalence tests (both halo paths) and the weak-scaling bench leg at 2 x 4M atoms.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k multi_gpu > $O/pytest_multi2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi2.log; tail -4 $O/pytest_multi2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 10 > $O/bench_multi_2.log 2>&1; echo "rc=$?" >> $O/bench_multi_2.log
tail -n 2 $O/bench_multi_2.log | cut -c1-1800
