#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_check.py 20 60 60 > gpurun_out/multi_check_$N.log 2>&1
echo "rc=$?" >> gpurun_out/multi_check_$N.log
tail -n 6 gpurun_out/multi_check_$N.log | cut -c1-1200
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k multi_gpu 2>&1 | tail -2
PISB_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_multi_$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_multi_$N.log
grep -E "pisb rank 0|rc=" gpurun_out/bench_multi_$N.log | tail -3 | cut -c1-300; tail -n 2 gpurun_out/bench_multi_$N.log | cut -c1-1500
