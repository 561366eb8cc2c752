#!/bin/bash
# round 2, visit G (2 GPUs): suite on the tree with the multi-type v3 build, proactive capacity growth, NVT on bricks; then the
# N = 1 bench as the driver runs it (new roofline layout) and the 2-GPU equivalence cases
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > $O/r02_pytest_g.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_g.log
tail -n 30 $O/r02_pytest_g.log
timeout 600 python bench.py > $O/r02_bench_g.log 2> $O/r02_bench_g.err; echo "bench rc=$?"
tail -c 2500 $O/r02_bench_g.log; tail -n 5 $O/r02_bench_g.err
