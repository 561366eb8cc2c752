#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_1gpu_4M.log 2> gpurun_out/bench_1gpu_4M.err; tail -c 4000 gpurun_out/bench_1gpu_4M.log; tail -3 gpurun_out/bench_1gpu_4M.err
timeout 900 python bench.py --steps 100 --warmup 10 --temperature 5 --no-cpu-baseline > gpurun_out/bench_1gpu_4M_5K.log 2>&1; tail -c 1500 gpurun_out/bench_1gpu_4M_5K.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; cat gpurun_out/launches_summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
