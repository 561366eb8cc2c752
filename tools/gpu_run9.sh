#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vv|k_copy_back|k_permute|k_sort_cells|k_fill|k_bin" -s 8 -c 8 -f -o gpurun_out/prof_stream python tools/prof_one.py 3 2 100 6 43 1 > gpurun_out/ncu_s.log 2>&1; tail -n 1 gpurun_out/ncu_s.log
