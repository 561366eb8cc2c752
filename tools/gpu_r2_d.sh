#!/bin/bash
# round 2, visit D: k_force_tile (shared-memory brick tiles): correctness, A/B against the thread-per-atom kernel, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_kernels.py tests/test_gpu_parity.py -m gpu -q -x -k "step_kernel or tile or four_lanes or compute_potential_parity or fused_force or unwrapped or pipelined or guard_band" > gpurun_out/r02_pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_d.log
tail -n 15 gpurun_out/r02_pytest_d.log
: > gpurun_out/r02_ab_tile.jsonl
for opt in "force_variant=0" "force_variant=8"; do
  timeout 300 python bench.py --steps 24 --warmup 6 --no-strong --no-cpu-baseline --e2e-steps 3 --option $opt >> gpurun_out/r02_ab_tile.jsonl 2>> gpurun_out/r02_ab_tile.err
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02_ab_tile.jsonl'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['options'], 'ms/step', round(d['ms_per_step'],4), d['kernel_ms_per_step'], 'builds', d['list_builds_in_timed_region'], 'drift', d['energy_drift_rel'])
PY
tail -n 3 gpurun_out/r02_ab_tile.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_force_tile" -s 30 -c 1 -f -o gpurun_out/r02_prof_force_tile python tools/prof_one.py 8 0 100 40 43 0 cuda_graphs=0 > gpurun_out/r02_ncu_force_tile.log 2>&1; tail -n 2 gpurun_out/r02_ncu_force_tile.log
