#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "npt" > gpurun_out/pytest_npt.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_npt.log
tail -n 60 gpurun_out/pytest_npt.log
