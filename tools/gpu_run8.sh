#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compute_potential_parity or bitwise" > gpurun_out/pytest_v4.log 2>&1; tail -n 5 gpurun_out/pytest_v4.log
for fv in 3 4; do timeout 300 python tools/variants_one.py $fv 2 1 100 43 40; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force_v4" -s 2 -c 1 -f -o gpurun_out/prof_force_v4 python tools/prof_one.py 4 2 100 3 43 1 > gpurun_out/ncu_v4.log 2>&1; tail -n 1 gpurun_out/ncu_v4.log
