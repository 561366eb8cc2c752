#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for mode in 2 1; do
PISB_HALO_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$mode tools/multi_check.py 16 60 60 2>&1 | grep -E "^\{" | cut -c1-400
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k multi_gpu 2>&1 | tail -2
