#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for V in 1 0; do
PISB_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$V bench.py --gpus 2 --steps 20 --warmup 10 --e2e-steps 60 --option fuse_vv=$V > $O/bench_multi_2_fuse$V.log 2>&1
grep "pisb rank 0" $O/bench_multi_2_fuse$V.log | tail -7 | cut -c1-200
python - <<PY
import json
for ln in open("$O/bench_multi_2_fuse$V.log"):
    if ln.startswith("{"):
        d=json.loads(ln); print("fuse_vv=$V value %.4g ms %.4f e2e %.4g ms %.4f builds %s/%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["list_builds_in_timed_region"], d.get("list_builds_in_e2e_region")))
PY
done
