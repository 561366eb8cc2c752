#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
timeout 600 python tools/variants.py 100 43 40 > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force|k_build_list" -s 3 -c 3 -f -o gpurun_out/prof_v2 python tools/prof_one.py 2 2 100 4 > gpurun_out/ncu_v2.log 2>&1; tail -n 3 gpurun_out/ncu_v2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force|k_build_list" -s 3 -c 3 -f -o gpurun_out/prof_v1 python tools/prof_one.py 1 1 100 4 > gpurun_out/ncu_v1.log 2>&1; tail -n 3 gpurun_out/ncu_v1.log
ls -la gpurun_out/
