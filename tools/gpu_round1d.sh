#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/fvv_variants.py 100 43 0,1 > $O/fvv_variants.jsonl 2> $O/fvv_variants.err; cat $O/fvv_variants.jsonl; tail -3 $O/fvv_variants.err
timeout 200 python tools/diag_e2e.py 100 > $O/diag_e2e.jsonl 2> $O/diag_e2e.err; cat $O/diag_e2e.jsonl; tail -3 $O/diag_e2e.err
for K in "k_force_vv fuse_vv=1 prof_force_vv" "k_force_v3 fuse_vv=0 prof_force_v3c"; do
  set -- $K
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 2 -c 1 -f -o $O/$3 python tools/prof_one.py 0 0 100 6 43 0 cuda_graphs=0 $2 > $O/ncu_$3.log 2>&1; tail -n 1 $O/ncu_$3.log
  python tools/ncu_summary.py $O/$3.ncu-rep > $O/$3.summary.json 2>>$O/ncu_summary.err
  ncu -i $O/$3.ncu-rep --page details > $O/$3.details.txt 2>/dev/null
  rm -f $O/$3.ncu-rep
done
