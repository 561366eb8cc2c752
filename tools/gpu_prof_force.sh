#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force_v3" -s 2 -c 1 -f -o gpurun_out/prof_force_v3 python tools/prof_one.py 3 3 100 3 43 0 > gpurun_out/ncu_f3.log 2>&1; tail -n 2 gpurun_out/ncu_f3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_build_list_v3" -s 0 -c 1 -f -o gpurun_out/prof_build_v3 python tools/prof_one.py 3 3 100 3 43 0 > gpurun_out/ncu_b3.log 2>&1; tail -n 2 gpurun_out/ncu_b3.log
