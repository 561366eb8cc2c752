#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_under_ncu.log 2>&1; tail -n 1 gpurun_out/bench_under_ncu.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force_v3" -s 6 -c 1 -f -o gpurun_out/prof_force_final python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_c.log 2>&1; tail -n 1 gpurun_out/ncu_c.log | cut -c1-200
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r01.log 2>&1; tail -n 1 gpurun_out/bench_r01.log
timeout 600 python bench.py --steps 100 --warmup 10 --temperature 5 --no-cpu-baseline > gpurun_out/bench_r01_5K.log 2>&1; tail -n 1 gpurun_out/bench_r01_5K.log | cut -c1-600
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -n 1 gpurun_out/bench_ref.log | cut -c1-300
