#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_1gpu_4M.log 2> gpurun_out/bench_1gpu_4M.err; tail -c 4000 gpurun_out/bench_1gpu_4M.log; tail -3 gpurun_out/bench_1gpu_4M.err
timeout 900 python bench.py --steps 100 --warmup 10 --temperature 5 --no-cpu-baseline > gpurun_out/bench_1gpu_4M_5K.log 2>&1; tail -c 1500 gpurun_out/bench_1gpu_4M_5K.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; cat gpurun_out/launches_summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force_v3" -s 2 -c 1 -f -o gpurun_out/prof_force_v3 python tools/prof_one.py 0 0 100 3 43 0 > gpurun_out/ncu_f3.log 2>&1; tail -n 1 gpurun_out/ncu_f3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_build_list_v3" -s 0 -c 1 -f -o gpurun_out/prof_build_v3 python tools/prof_one.py 0 0 100 3 43 0 > gpurun_out/ncu_b3.log 2>&1; tail -n 1 gpurun_out/ncu_b3.log
