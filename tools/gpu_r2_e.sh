#!/bin/bash
# round 2, visit E: PROBE -- how fast is the step kernel when 2 / 4 adjacent lanes gather the same neighbours (same-address
# merge in L1TEX)?  (wrong physics on purpose, timing only); then the whole -m gpu suite on the cleaned-up tree
mkdir -p gpurun_out
: > gpurun_out/r02_probe_share.jsonl
for p in 0 1 3; do
  PISB_SHARE_PROBE=$p timeout 300 python bench.py --steps 12 --warmup 4 --no-strong --no-cpu-baseline --e2e-steps 3 >> gpurun_out/r02_probe_share.jsonl 2>> gpurun_out/r02_probe_share.err
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02_probe_share.jsonl'):
    if ln.startswith('{'):
        d=json.loads(ln); print('ms/step', round(d['ms_per_step'],4), d['kernel_ms_per_step'], 'f_ms', round(d['roofline']['ms_per_launch'],4), 'builds', d['list_builds_in_timed_region'])
PY
tail -n 3 gpurun_out/r02_probe_share.err
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_e.log
tail -n 25 gpurun_out/r02_pytest_e.log
