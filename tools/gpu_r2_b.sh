#!/bin/bash
# round 2, visit B: suite on the four-lanes-per-atom default, A/B of the step kernels at 4M atoms / 43 K, ncu of the new default
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_b.log
tail -n 25 gpurun_out/r02_pytest_b.log
: > gpurun_out/r02_ab_step_kernels.jsonl
for opt in "force_variant=0" "force_variant=3" "force_variant=5"; do
  timeout 300 python bench.py --steps 24 --warmup 6 --no-strong --no-cpu-baseline --e2e-steps 3 --option $opt >> gpurun_out/r02_ab_step_kernels.jsonl 2>> gpurun_out/r02_ab.err
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02_ab_step_kernels.jsonl'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['options'], 'ms/step', round(d['ms_per_step'],4), d['kernel_ms_per_step'], 'f_ms', round(d['roofline']['ms_per_launch'],4), 'builds', d['list_builds_in_timed_region'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_force_q" -s 30 -c 1 -f -o gpurun_out/r02_prof_force_q python tools/prof_one.py 0 0 100 40 43 0 cuda_graphs=0 > gpurun_out/r02_ncu_force_q.log 2>&1; tail -n 2 gpurun_out/r02_ncu_force_q.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_build_list_v3" -s 4 -c 1 -f -o gpurun_out/r02_prof_build_v3 python tools/prof_one.py 0 0 100 40 43 0 cuda_graphs=0 > gpurun_out/r02_ncu_build_v3.log 2>&1; tail -n 2 gpurun_out/r02_ncu_build_v3.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke_b.log 2>&1; tail -n 3 gpurun_out/r02_smoke_b.log
