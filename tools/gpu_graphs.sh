#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "graph or npt or hot_run or nve_trace" > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_graph.log
tail -n 40 gpurun_out/pytest_graph.log
timeout 600 python tools/small_systems.py > gpurun_out/small_systems.jsonl 2> gpurun_out/small_systems.err; cat gpurun_out/small_systems.jsonl; tail -5 gpurun_out/small_systems.err
