#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/graph_cond_probe > gpurun_out/graph_probe.log 2>&1; echo "rc=$?" >> gpurun_out/graph_probe.log
cat gpurun_out/graph_probe.log
timeout 600 python -m pytest tests -m gpu -q -k "npt" > gpurun_out/pytest_npt.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_npt.log
tail -n 30 gpurun_out/pytest_npt.log
