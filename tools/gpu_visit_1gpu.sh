#!/bin/bash
# single-GPU visit: the whole -m gpu suite on the final tree, smoke(), bench.py as the driver runs it, the reference
# arm, the ncu launch list of a short bench and the --set full capture of the step kernel, memcheck / racecheck of smoke()
mkdir -p gpurun_out
O=gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; git rev-parse --short HEAD 2>/dev/null; } > $O/r02_box_final.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
tail -n 16 $O/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log; tail -n 4 $O/r02_smoke.log
timeout 600 python bench.py > $O/r02_bench_1gpu_4M.log 2> $O/r02_bench_1gpu_4M.err; echo "bench rc=$?"
tail -c 1500 $O/r02_bench_1gpu_4M.log; tail -n 3 $O/r02_bench_1gpu_4M.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference_arm.log 2> $O/r02_bench_reference_arm.err; echo "reference rc=$?"
tail -c 800 $O/r02_bench_reference_arm.log
# launch list of the bench command with classic launches (kernel nodes of graphs with conditional nodes cannot be listed): warm-up,
# timed and profiled legs dominate the list, so the kernel shares are comparable with the bench line's kernel_ms_per_step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches.csv python bench.py --steps 100 --warmup 10 --no-strong --no-cpu-baseline --e2e-steps 3 --option cuda_graphs=0 > $O/r02_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_force_vv" -s 30 -c 1 -f -o $O/r02_prof_force_vv_final python tools/prof_one.py 0 0 100 40 43 0 cuda_graphs=0 > $O/r02_ncu_force_vv_final.log 2>&1; tail -n 2 $O/r02_ncu_force_vv_final.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py --smoke > $O/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/r02_sanitizer_memcheck.log
# racecheck: the small-system leg of smoke() (graph replays), then the kernels of the 4M-atom benchmark at 108 000 atoms with classic
# launches (with graph replays of that size the tool's target process dies -- not a hazard report; memcheck above covers that run)
PISB_SMOKE_SMALL=1 timeout 420 compute-sanitizer --tool racecheck --error-exitcode 3 python __graft_entry__.py --smoke > $O/r02_sanitizer_racecheck.log 2>&1; echo "racecheck (smoke, 2048 atoms) rc=$?" | tee -a $O/r02_sanitizer_racecheck.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_large.py 0 >> $O/r02_sanitizer_racecheck.log 2>&1; echo "racecheck (108000 atoms, classic launches) rc=$?" | tee -a $O/r02_sanitizer_racecheck.log
tail -n 3 $O/r02_sanitizer_memcheck.log $O/r02_sanitizer_racecheck.log
timeout 300 python tools/small_systems.py > $O/r02_small_systems.jsonl 2> $O/r02_small_systems.err; echo "small rc=$?"; tail -n 12 $O/r02_small_systems.jsonl | cut -c1-400
