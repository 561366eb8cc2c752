#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "graph or bench" > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_graph.log; tail -4 gpurun_out/pytest_graph.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_1gpu_4M.log 2> gpurun_out/bench_1gpu_4M.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu_4M.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","ms_per_step_profiled","gpu_launches","kernel_ms_per_step","list_builds_in_timed_region","list_builds_in_profiled_pass")}, d["e2e"]["value"], d["e2e_resident"]["value"])
PY
timeout 600 python tools/small_systems.py > gpurun_out/small_systems.jsonl 2> gpurun_out/small_systems.err; cat gpurun_out/small_systems.jsonl
