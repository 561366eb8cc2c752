#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/variants_lib.log
for lib in build/variants/*.so; do
  echo "== $lib" >> gpurun_out/variants_lib.log
  PISB_LIB=$PWD/$lib timeout 300 python tools/variants_one.py ${FV:-3} ${BV:-2} ${CD:-1} >> gpurun_out/variants_lib.log 2>&1
  if [ -n "$CD2" ]; then PISB_LIB=$PWD/$lib timeout 300 python tools/variants_one.py ${FV:-3} ${BV:-2} 2 >> gpurun_out/variants_lib.log 2>&1; fi
done
grep -E "^==|ms_per_step|Error|error" gpurun_out/variants_lib.log | sed -E 's/\{"fv".*"cd": ([0-9]), "ms_per_step": ([0-9.]+).*"build": ([0-9.]+), "force": ([0-9.]+).*"builds": ([0-9]+).*/cd \1 step \2 build \3 force \4 builds \5/'
