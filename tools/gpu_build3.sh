#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "neighbour_list or compute_potential_parity or prefilter or guard_band or multi_gpu or golden" > gpurun_out/pytest_b3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b3.log
tail -n 25 gpurun_out/pytest_b3.log
timeout 600 python tools/variants.py 100 43.0 60 > gpurun_out/variants_b3.jsonl 2> gpurun_out/variants_b3.err; cat gpurun_out/variants_b3.jsonl; tail -3 gpurun_out/variants_b3.err
