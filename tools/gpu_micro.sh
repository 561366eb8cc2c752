#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/microbench > gpurun_out/microbench_pipes.txt 2>&1; cat gpurun_out/microbench_pipes.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_4000.csv python tools/prof_small.py > gpurun_out/prof_small.log 2>&1; tail -3 gpurun_out/prof_small.log
python tools/launch_summary.py gpurun_out/launches_4000.csv | tee gpurun_out/launches_4000_summary.txt
