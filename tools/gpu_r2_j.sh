#!/bin/bash
# round 2, visit J (2 GPUs): why the NVT leg of multi_parity reported KE = NaN at 8 GPUs (single-GPU half alone, then the
# bench's check at 2 GPUs), and the k_force_vv block-size / SM-aware scheduling A/B at 4M atoms (1 GPU)
mkdir -p gpurun_out
O=gpurun_out
for a in "24 40 3" "24 40 0" "16 60 3"; do timeout 200 python tools/diag_nvt.py $a; done > $O/r02_diag_nvt.jsonl 2> $O/r02_diag_nvt.err
cat $O/r02_diag_nvt.jsonl | cut -c1-900; tail -n 3 $O/r02_diag_nvt.err
PISB_FORCE_VARIANT=3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_check.py 24 40 > $O/r02_multi_check_2gpu.log 2>&1; echo "rc=$?" >> $O/r02_multi_check_2gpu.log
tail -n 3 $O/r02_multi_check_2gpu.log | cut -c1-2500
timeout 900 python tools/force_sched_ab.py 100 43 48 > $O/r02_force_sched_ab.jsonl 2> $O/r02_force_sched_ab.err; echo "ab rc=$?"
cat $O/r02_force_sched_ab.jsonl; tail -n 5 $O/r02_force_sched_ab.err
