#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/build_variants.py 100 8,6,10,12 > $O/build_variants.jsonl 2> $O/build_variants.err; cat $O/build_variants.jsonl; tail -3 $O/build_variants.err
