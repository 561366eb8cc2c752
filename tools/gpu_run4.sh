#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
timeout 600 python tools/variants.py 100 43 40 > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force" -s 2 -c 1 -f -o gpurun_out/prof_force_v3 python tools/prof_one.py 3 2 100 3 > gpurun_out/ncu_a.log 2>&1; tail -n 2 gpurun_out/ncu_a.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_build_list" -c 1 -f -o gpurun_out/prof_build_v2 python tools/prof_one.py 3 2 100 1 > gpurun_out/ncu_b.log 2>&1; tail -n 2 gpurun_out/ncu_b.log
