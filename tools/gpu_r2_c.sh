#!/bin/bash
# round 2, visit C: texture-pipe gather experiment, step-kernel choice across system sizes, ncu of the default step kernel + a real build
mkdir -p gpurun_out
: > gpurun_out/r02_ab_tex.jsonl
for opt in "tex_gather=0" "tex_gather=1" "tex_gather=2"; do
  timeout 300 python bench.py --steps 24 --warmup 6 --no-strong --no-cpu-baseline --e2e-steps 3 --option $opt >> gpurun_out/r02_ab_tex.jsonl 2>> gpurun_out/r02_ab_tex.err
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02_ab_tex.jsonl'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['options'], 'ms/step', round(d['ms_per_step'],4), d['kernel_ms_per_step'], 'builds', d['list_builds_in_timed_region'], 'drift', d['energy_drift_rel'])
PY
timeout 600 python tools/small_systems.py > gpurun_out/r02_small_systems.jsonl 2> gpurun_out/r02_small_systems.err; cat gpurun_out/r02_small_systems.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    d=json.loads(ln); print(d['n_atoms'], d['T0'], 'fv', d['force_variant'], d['us_per_step'], 'us/step')"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_force_vv" -s 30 -c 1 -f -o gpurun_out/r02_prof_force_vv python tools/prof_one.py 0 0 100 40 43 0 cuda_graphs=0 > gpurun_out/r02_ncu_force_vv.log 2>&1; tail -n 2 gpurun_out/r02_ncu_force_vv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_build_list_v3" -s 42 -c 1 -f -o gpurun_out/r02_prof_build_v3 python tools/prof_one.py 0 0 100 40 43 0 cuda_graphs=0 > gpurun_out/r02_ncu_build_v3.log 2>&1; tail -n 2 gpurun_out/r02_ncu_build_v3.log
timeout 300 python -m pytest tests/test_gpu_quad.py tests/test_gpu_parity.py -q -k "quad or odd or families or fused_force" > gpurun_out/r02_pytest_c.log 2>&1; tail -n 3 gpurun_out/r02_pytest_c.log
