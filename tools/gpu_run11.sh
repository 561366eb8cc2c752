#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vv" -s 3 -c 2 -f -o gpurun_out/prof_vv python tools/prof_one.py 3 2 100 6 5 1 > gpurun_out/ncu_vv.log 2>&1; tail -n 1 gpurun_out/ncu_vv.log
