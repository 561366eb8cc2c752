#!/bin/bash
# strong scaling of the 32M-atom system: N = 1 (single-GPU path) or N > 1 (torchrun, --strong)
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --ncell 200 --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_strong_1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_strong_1.log
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --strong --steps 30 --warmup 5 > gpurun_out/bench_strong_$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_strong_$N.log
fi
grep -E "^\{|rc=" gpurun_out/bench_strong_$N.log | cut -c1-700
