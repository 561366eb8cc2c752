#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 50 python -m pytest tests -m gpu -q -n 8 -x > $O/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_final.log
tail -n 4 $O/pytest_gpu_final.log
timeout 16 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 5 > $O/bench_1gpu_4M_final_short.log 2>&1; tail -c 1500 $O/bench_1gpu_4M_final_short.log
