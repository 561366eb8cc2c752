#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "nvt or cli" > gpurun_out/pytest_nvt.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_nvt.log
tail -n 25 gpurun_out/pytest_nvt.log
