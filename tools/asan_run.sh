#!/bin/bash
# Host-side AddressSanitizer pass over the library (tools/asan/libpisb200_asan.so, built by the command in tools/README.md):
# the graph-replay step loop at 108 000 atoms, then the graph / fused-step / NVT / NPT / host-step cases of the GPU suite.
mkdir -p gpurun_out
O=gpurun_out
export LD_PRELOAD=$(gcc -print-file-name=libasan.so)
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0:halt_on_error=0
export PISB_LIB=$PWD/tools/asan/libpisb200_asan.so
timeout 300 python tools/sanitize_large.py 1 > $O/r02_asan_large.log 2>&1; echo "asan large rc=$?" | tee -a $O/r02_asan_large.log
tail -n 40 $O/r02_asan_large.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graph or fused or nvt or npt or pipelined or capacity or velocity or asynchronous" > $O/r02_asan_pytest.log 2>&1; echo "asan pytest rc=$?" | tee -a $O/r02_asan_pytest.log
grep -n "ERROR: AddressSanitizer\|SUMMARY: AddressSanitizer\|passed\|failed" $O/r02_asan_pytest.log | head -20
