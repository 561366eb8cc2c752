#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_gpu.log
timeout 600 python tools/variants.py 100 43 40 > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
timeout 600 python tools/variants.py 100 5 40 > gpurun_out/variants_5K.log 2>&1; cat gpurun_out/variants_5K.log
