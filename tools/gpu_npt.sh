#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "npt" > gpurun_out/pytest_npt.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_npt.log
tail -n 60 gpurun_out/pytest_npt.log
timeout 1500 python -m pytest tests -m gpu -x -q -k "not npt" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
