#!/bin/bash
# One visit of the B200 box: the -m gpu suite (6 workers), ncu --set full captures of the step kernel and the list build
# (classic launches: ncu cannot profile kernel nodes of graphs that hold conditional nodes) -> force_traffic.json,
# bench.py at 4M atoms (43 K with the CPU baseline, 5 K, and 43 K with the fused step kernel / host pipeline off),
# the ncu launch list of a short bench run, small systems, smoke().
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -n 6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -n 6 $O/pytest_gpu.log
for K in "k_force_vv 2 prof_force_vv" "k_build_list_v3 0 prof_build_v3"; do
  set -- $K
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c 1 -f -o $O/$3 python tools/prof_one.py 0 0 100 8 43 0 cuda_graphs=0 > $O/ncu_$3.log 2>&1; tail -n 1 $O/ncu_$3.log
  python tools/ncu_summary.py $O/$3.ncu-rep > $O/$3.summary.json 2>>$O/ncu_summary.err
  ncu -i $O/$3.ncu-rep --page details > $O/$3.details.txt 2>/dev/null
  rm -f $O/$3.ncu-rep
done
python tools/force_traffic.py $O/prof_force_vv.summary.json profiles/force_traffic.json "profiles/r01_force_vv_full.summary.json (ncu --set full, 4M atoms, T0=43K, classic launches)" > /dev/null && cp profiles/force_traffic.json $O/force_traffic.json
timeout 300 python bench.py --steps 100 --warmup 10 > $O/bench_1gpu_4M.log 2> $O/bench_1gpu_4M.err; tail -c 1200 $O/bench_1gpu_4M.log; tail -3 $O/bench_1gpu_4M.err
timeout 200 python bench.py --steps 100 --warmup 10 --temperature 5 --no-cpu-baseline > $O/bench_1gpu_4M_5K.log 2>&1; tail -c 300 $O/bench_1gpu_4M_5K.log
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --option fuse_vv=0 --option host_pipeline=0 > $O/bench_1gpu_4M_unfused.log 2>&1; tail -c 300 $O/bench_1gpu_4M_unfused.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 3 > $O/bench_under_ncu.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launches_summary.txt; cat $O/launches_summary.txt
timeout 200 python tools/small_systems.py > $O/small_systems.jsonl 2>$O/small_systems.err; cat $O/small_systems.jsonl | cut -c1-220
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.log 2>&1; tail -c 400 $O/bench_reference_arm.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
