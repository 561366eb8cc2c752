#!/bin/bash
# One bounded GPU visit: full -m gpu suite (4 workers), ncu capture of the fused step kernel -> force_traffic.json,
# bench with the defaults (fused step kernel, pipelined host step) and with both switched off, launch list, smoke.
mkdir -p gpurun_out
O=gpurun_out
timeout 480 python -m pytest tests -m gpu -q -n 6 --durations=12 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 30 $O/pytest_gpu.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_force_vv" -s 1 -c 1 -f -o $O/prof_force_vv python tools/prof_one.py 0 0 100 6 43 0 > $O/ncu_fvv.log 2>&1; tail -n 2 $O/ncu_fvv.log
python tools/ncu_summary.py $O/prof_force_vv.ncu-rep > $O/prof_force_vv.summary.json 2>$O/ncu_summary.err
ncu -i $O/prof_force_vv.ncu-rep --page details > $O/prof_force_vv.details.txt 2>/dev/null
python tools/force_traffic.py $O/prof_force_vv.summary.json profiles/force_traffic.json "profiles/r01_force_vv_full.summary.json (ncu --set full, 4M atoms, T0=43K)" && cp profiles/force_traffic.json $O/force_traffic.json
timeout 300 python bench.py --steps 100 --warmup 10 > $O/bench_1gpu_4M.log 2> $O/bench_1gpu_4M.err; tail -c 3500 $O/bench_1gpu_4M.log; tail -3 $O/bench_1gpu_4M.err
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --option fuse_vv=0 --option host_pipeline=0 > $O/bench_1gpu_4M_unfused.log 2>&1; tail -c 3000 $O/bench_1gpu_4M_unfused.log
timeout 200 python bench.py --steps 100 --warmup 10 --temperature 5 --no-cpu-baseline > $O/bench_1gpu_4M_5K.log 2>&1; tail -c 600 $O/bench_1gpu_4M_5K.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 3 > $O/bench_under_ncu.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launches_summary.txt; cat $O/launches_summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
