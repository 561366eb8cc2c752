#!/bin/bash
# first GPU session: parity tests, smoke, microbench, first bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 120 ./tools/microbench > gpurun_out/microbench.log 2>&1
timeout 900 python bench.py --steps 50 --warmup 10 --cpu-steps 3 > gpurun_out/bench1.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench1.log
tail -n 25 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log; cat gpurun_out/microbench.log; tail -n 4 gpurun_out/bench1.log
