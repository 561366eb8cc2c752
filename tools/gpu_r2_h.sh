#!/bin/bash
# round 2, visit H (1 GPU): list-layout tests (build window, aligned rows), the multi-type test that failed in visit G (hot run
# now at a stable dt), then the layout A/B at 4M atoms and the rest of the suite
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_list_layout.py tests/test_gpu_parity.py -m gpu -q -x -k "list_layout or multi_type or neighbour_list_exact or compute_potential_parity" > $O/r02_pytest_h1.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_h1.log
tail -n 15 $O/r02_pytest_h1.log
timeout 600 python tools/layout_ab.py 100 43 48 > $O/r02_layout_ab.jsonl 2> $O/r02_layout_ab.err; echo "ab rc=$?"
cat $O/r02_layout_ab.jsonl; tail -n 5 $O/r02_layout_ab.err
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > $O/r02_pytest_h2.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_h2.log
tail -n 12 $O/r02_pytest_h2.log
