#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python tools/small_systems.py > gpurun_out/small_systems.jsonl 2> gpurun_out/small_systems.err; head -8 gpurun_out/small_systems.jsonl
