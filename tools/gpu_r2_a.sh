#!/bin/bash
# round 2, visit A: the whole -m gpu suite (incl. the full-size configuration tests), a short bench, sanitizer logs
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; } > gpurun_out/r02_box_a.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02_pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_a.log
tail -n 45 gpurun_out/r02_pytest_a.log
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/r02_bench_a.log 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench_a.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py --smoke > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 3 python __graft_entry__.py --smoke > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck.log
tail -n 5 gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log
