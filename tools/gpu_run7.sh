#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/small_systems.py > gpurun_out/small_systems.log 2>&1; cat gpurun_out/small_systems.log
timeout 1500 python tools/sweep_c5.py 100 20 43 > gpurun_out/sweep_c5.log 2>&1; cat gpurun_out/sweep_c5.log | cut -c1-420
