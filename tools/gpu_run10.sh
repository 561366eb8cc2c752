#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "parity or exact or bitwise or rebuilds or golden" > gpurun_out/pytest_q.log 2>&1; tail -n 3 gpurun_out/pytest_q.log
timeout 300 python tools/variants_one.py 3 2 1 100 43 40
timeout 300 python tools/variants_one.py 3 2 1 100 5 40
