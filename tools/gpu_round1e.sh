#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -n 6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 12 $O/pytest_gpu.log
timeout 200 python tools/fvv_variants.py 100 43 0,1,2 > $O/fvv_variants2.jsonl 2> $O/fvv_variants2.err; cat $O/fvv_variants2.jsonl; tail -3 $O/fvv_variants2.err
timeout 100 python tools/fvv_variants.py 100 5 0,1 > $O/fvv_variants2_5K.jsonl 2> $O/fvv_variants2.err; cat $O/fvv_variants2_5K.jsonl
timeout 100 python tools/fvv_variants.py 40 43 0,1 > $O/fvv_variants2_256k.jsonl 2> $O/fvv_variants2.err; cat $O/fvv_variants2_256k.jsonl
