#!/usr/bin/env python
"""bench.py -- atom-steps/s of the LJ argon NVE hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A "step" is one NVE timestep (drift + wrap + skin check + [list rebuild] + LJ force + kick + thermo)
of the whole synthetic FCC-argon system.  N=1 runs BASELINE configs[2] (4M atoms, rc = 2.5 sigma); N>1 runs the spatially decomposed path
with one 4M-atom brick per GPU (weak scaling: N=8 is configs[3], 32M atoms on 2x2x2 bricks);
--strong runs the 32M-atom system on any N.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "atom-steps/sec LJ argon NVE"
UNIT = "atom-steps/s"
DT = 0.25          # example/input.pis:8
SIGMA = 3.405
RC = 2.5 * SIGMA   # BASELINE configs: rc = 2.5 sigma
SKIN = 0.3 * SIGMA


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML through pynvml when it is there
    (a query takes tens of microseconds: one sample every 5 ms, so even a 30 ms region holds several); the nvidia-smi command
    line of the recipe otherwise (one sample per ~0.2 s)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.source = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            try:  # the CUDA device, not the NVML index (CUDA_VISIBLE_DEVICES may renumber)
                import torch
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(index).uuid)).encode())
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = (pynvml, handle)
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv, hd = self._nvml
        sm = nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(hd, nv.NVML_CLOCK_SM)
        try:
            pw = nv.nvmlDeviceGetPowerUsage(hd) / 1000.0
        except Exception:
            pw = "n/a"
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(hd) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(hd)
        act = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(sm), str(mx), str(pw), act(nv.nvmlClocksEventReasonHwSlowdown), act(nv.nvmlClocksEventReasonHwThermalSlowdown),
                act(nv.nvmlClocksEventReasonSwThermalSlowdown), act(nv.nvmlClocksEventReasonSwPowerCap)]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    parts = [p.strip() for p in out.strip().split(",")]
                    if len(parts) >= 7:
                        self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.005 if self._nvml else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        try:
            pw = max(float(r[2]) for r in self.rows)
        except ValueError:
            pw = None
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "power_w_max": pw,
                "samples": len(self.rows), "source": self.source, "reasons": reasons}


def measure_pcie(host_array, reps: int = 4) -> dict:
    """Host<->device copy rates of one of the run's own pinned 3N-double arrays, each direction alone (the array is
    overwritten: the caller downloads the state into it afterwards)."""
    import torch

    host = torch.from_numpy(host_array.reshape(-1))
    n_doubles = host.numel()
    dev = torch.empty(n_doubles, dtype=torch.float64, device="cuda")
    out = {}
    for name, fn in (("h2d_GBps", lambda: dev.copy_(host, non_blocking=True)), ("d2h_GBps", lambda: host.copy_(dev, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        out[name] = 8.0 * n_doubles * reps / (time.perf_counter() - t0) / 1e9
    return out


# FP64 yardsticks of the force loop (the resource that binds it; the HBM view is reported beside it).
#  * ALGORITHMIC flops per atom-step (full list, div and sqrt counted as 1):
#        21 per LISTED pair (difference 3, minimum image 12, r2 5, compare 1) + 18 per IN-RANGE pair + ~20 (integrator)
#    -> 21 K + 18 K_in + 20.  K and K_in are measured on the run's own list (pisb_pairs_in_range).
#  * ISSUED slots per listed pair of the lean loop, counted in the SASS: 6 distance + 4 Newton
#    reciprocal + 6 LJ + 5 accumulate + 2 loop overhead = 23 (round 1: 35); a warp pays the in-range part whenever ANY of
#    its lanes is in range, so issue is charged per listed pair.  The fused kernel's epilogue adds ~110 per atom.
#  * Peak: the measured DFMA issue rate of tools/microbench.cu, 57.16 per clock and SM (profiles/r01_microbench_pipes.txt),
#    148 SMs at 1965 MHz = 16.6 T instr/s, i.e. 33.2 TFLOP/s when every slot is a fused multiply-add.  (There is no
#    driver-written FP64 peak in MEASURED_PEAKS.json.)
FP64_SLOTS_PER_PAIR = 23.0
FP64_SLOTS_EPILOGUE = 110.0
FP64_PEAK_SLOTS_PER_S = 57.16 * 148 * 1.965e9


def fp64_roofline(kernel: str, k_mean: float, k_in: float, n_atoms: int, ms_per_launch: float, fused: bool) -> dict:
    flops = 21.0 * k_mean + 18.0 * k_in + (20.0 if fused else 0.0)
    slots = FP64_SLOTS_PER_PAIR * k_mean + (FP64_SLOTS_EPILOGUE if fused else 0.0)
    t = ms_per_launch * 1e-3
    achieved = flops * n_atoms / t
    return {"kernel": kernel, "bound": "fp64", "achieved": achieved / 1e12, "peak": 2.0 * FP64_PEAK_SLOTS_PER_S / 1e12, "unit": "TFLOP/s",
            "frac": achieved / (2.0 * FP64_PEAK_SLOTS_PER_S),
            "algorithmic_flops_per_atom": flops, "mean_neighbours_listed": k_mean, "mean_neighbours_in_range": k_in,
            "frac_of_dfma_issue_rate": achieved / FP64_PEAK_SLOTS_PER_S,
            "issued_fp64_slots_per_atom": slots, "issue_frac": slots * n_atoms / t / FP64_PEAK_SLOTS_PER_S,
            "peak_source": "measured DFMA issue rate x 2 flop (tools/microbench.cu: 57.16 per clock and SM, 148 SMs, 1965 MHz)",
            "ms_per_launch": ms_per_launch,
            "note": "algorithmic flops = 21 K + 18 K_in + 20 (SURVEY 8d); frac = algorithmic flops / (2 x measured DFMA rate); "
                    "frac_of_dfma_issue_rate = the same flops / the DFMA rate (one flop per slot); issue_frac = FP64 instructions "
                    "the kernel issues (SASS count x listed pairs) / the DFMA rate = the pipe's utilisation"}


# What actually binds the step kernel (ncu, profiles/r02_force_vv_final.*): the L1TEX DATA PIPE.  Every listed pair is one
# scattered 32-byte record, and the data stage delivers one 128-byte line ("wavefront") per clock and SM whatever part of the
# line is wanted: l1tex__data_pipe_lsu_wavefronts 89 % of peak, 0.83 wavefronts per listed pair (neighbours that share a
# line share a wavefront), while the FP64 pipe is 45 % busy and DRAM 24 %.  Live part: listed pairs per second from this run's
# list and launch time; the wavefronts-per-pair factor is the static ncu calibration (profiles/force_traffic.json).
def l1tex_roofline(kernel: str, k_mean: float, n_atoms: int, ms_per_launch: float, sm_mhz, cal: dict) -> dict:
    clock = float(sm_mhz or 1965.0) * 1e6
    pairs_per_s = k_mean * n_atoms / (ms_per_launch * 1e-3)
    peak = 148.0 * clock                                    # one data-pipe wavefront per clock and SM
    w = float(cal.get("l1tex_wavefronts_per_listed_pair") or 1.0)
    return {"kernel": kernel, "bound": "l1tex", "achieved": pairs_per_s * w / 1e9, "peak": peak / 1e9, "unit": "G wavefronts/s",
            "frac": pairs_per_s * w / peak, "listed_pairs_per_s": pairs_per_s, "wavefronts_per_listed_pair_ncu": w,
            "data_pipe_pct_ncu": cal.get("l1tex_data_pipe_lsu_wavefronts_pct"), "lsu_writeback_pct_ncu": cal.get("l1tex_lsu_writeback_active_pct"),
            "sm_clock_mhz": clock / 1e6, "ms_per_launch": ms_per_launch, "peak_source": "1 data-pipe wavefront per clock and SM x 148 SMs x the SM clock sampled under load",
            "note": f"THE BINDING RESOURCE of the force loop (ncu: l1tex__data_pipe_lsu_wavefronts {cal.get('l1tex_data_pipe_lsu_wavefronts_pct', float('nan')):.1f} % "
                    f"of peak; FP64 pipe {cal.get('fp64_pipe_active_pct', float('nan')):.0f} %, DRAM {cal.get('dram_throughput_pct', float('nan')):.0f} %): every listed pair "
                    "is a scattered 32-byte record and costs one 128-byte data-pipe wavefront.  achieved = live listed pairs/s x the "
                    "static ncu wavefronts-per-pair factor; *_ncu fields are static (profiles/force_traffic.json)"}


def build_roofline(tim: dict, builds: int, n_atoms: int, k_mean: float):
    """The second kernel of the step, k_build_list_v3 (a quarter of the step at 43 K): live ms per build from this run's per-kernel
    pass; what binds it comes from the static ncu capture profiles/r02_build_v3_full.summary.json -- instruction issue (issue
    slots ~69 % active, ALU pipe 58 %, 23 of 32 lanes), not memory (DRAM 15 %), so it is reported against the issue rate."""
    try:
        ms_per_build = tim["build"]["ms"] / max(builds, 1)
        out = {"kernel": "k_build_list_v3", "bound": "issue", "builds_in_profiled_pass": builds, "ms_per_build": ms_per_build,
               "listed_pairs_per_s": k_mean * n_atoms / (ms_per_build * 1e-3) if builds else None,
               "algorithmic_list_bytes_per_build": 4.0 * k_mean * n_atoms}
        sp = os.path.join(ROOT, "profiles", "r02_build_v3_full.summary.json")
        with open(sp) as f:
            rows = json.load(f)
        r = rows[0]
        pct = lambda key: float(str(r.get(key, "nan")).split()[0])
        out.update({"achieved": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"), "peak": 100.0, "unit": "% of issue slots (ncu, static)",
                    "frac": pct("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
                    "alu_pipe_pct_ncu": pct("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                    "dram_throughput_pct_ncu": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    "lanes_active_per_instruction_ncu": pct("smsp__thread_inst_executed_per_inst_executed.ratio"),
                    "warp_instructions_per_build_ncu": pct("smsp__inst_executed.sum"),
                    "source": "static ncu --set full capture profiles/r02_build_v3_full.summary.json; ms_per_build is live"})
        return out
    except Exception as e:  # the headline line must not be lost to an extra view
        return {"kernel": "k_build_list_v3", "error": repr(e)[:200]}


def force_calibration(kernel_name: str):
    """The static ncu capture of the dominant kernel (profiles/force_traffic.json): (all fields, dram bytes per launch, FP64 pipe
    %, provenance note) -- empty when the capture is of another kernel."""
    tp = os.path.join(ROOT, "profiles", "force_traffic.json")
    try:
        with open(tp) as f:
            tj = json.load(f)
    except Exception:
        return {}, None, None, None
    if not str(tj.get("kernel", "")).startswith(kernel_name.split("<")[0]):    # a capture of another kernel says nothing about this one
        return {}, None, None, None
    src = f"static ncu --set full capture {tj.get('source', 'profiles/force_traffic.json')} ({tj.get('commit', 'commit n/a')}); not re-measured in this run"
    return tj, tj.get("dram_bytes_per_launch"), tj.get("fp64_pipe_active_pct"), src


def argon_oracle(atoms):
    from oracle.pis_oracle import Oracle

    orc = Oracle(atoms.sim_box.h_colmajor(), masses=atoms.masses)
    orc.insert(1, 1, 0.238, SIGMA, RC)
    return orc


def host_cores() -> int:
    """Threads the CPU arm may use: the cores this process is allowed on.  NOT omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which silently made round 1's N > 1 reference arm a single-thread run."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_cpu_path(ncell: int, temperature: float, steps: int, warmup: int, threads: int, seed: int = 12345, budget_s: float = 0.0):
    """The reference's CPU path (oracle port, OpenMP all-core variant of the LJVOffsetManager loop) on an
    FCC-argon system of ncell^3*4 atoms: returns (atom-steps/s, cores, seconds per step list).  budget_s > 0: the first
    warm-up step is timed and, if (warmup + steps) of them would not fit the budget, None is returned instead (the caller
    picks a smaller sample)."""
    from pis_b200.lattice import fcc_argon

    atoms = fcc_argon(ncell, temperature=temperature, seed=seed)
    orc = argon_oracle(atoms)
    cores = threads if threads > 0 else host_cores()
    mode = "omp" if cores > 1 else "serial"
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    f = np.zeros_like(x)
    orc.compute_potential(x, atoms.type_ids, forces=f, mode=mode, threads=cores)
    for w in range(warmup):
        t0 = time.perf_counter()
        orc.verlet_step_nve(x, v, f, atoms.type_ids, DT, mode=mode, threads=cores)
        if w == 0 and budget_s > 0.0 and (time.perf_counter() - t0) * (warmup + steps) > budget_s:
            return None
    per = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.verlet_step_nve(x, v, f, atoms.type_ids, DT, mode=mode, threads=cores)
        per.append(time.perf_counter() - t0)
    tot = sum(per)
    return atoms.n_atoms * steps / tot, cores, per


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, on the box's host cores.  The Rust reference
    cannot be built here (no rustc), so this is the oracle port (kind "port"), all host threads.  At N = 1 it runs the REAL
    configuration it prints (4M atoms, K timed steps after W warm-up steps); the per-GPU block of the same lattice is the
    bounded sample at N > 1 (atom-steps/s is intensive), and a 256k-atom block replaces either if K + W steps of it would
    not end within a few minutes on this host -- the line says which was timed (`sample_atoms`, `cpu_baseline.sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cfg = workload_config(args, args.gpus)
    res, ncell = None, args.ncell
    if args.ref_ncell <= 0:                      # automatic: the real per-GPU block unless it blows the time budget
        res = time_cpu_path(ncell, args.temperature, args.steps, args.warmup, args.cpu_threads, budget_s=args.ref_budget_s)
    if res is None:
        ncell = args.ref_ncell if args.ref_ncell > 0 else 40
        res = time_cpu_path(ncell, args.temperature, args.steps, args.warmup, args.cpu_threads)
    value, cores, per = res
    n = 4 * ncell ** 3
    ms = 1e3 * sum(per) / len(per)
    whole = n == cfg["n_atoms"]
    sample = ((f"the whole {n}-atom workload" if whole else
               f"a {n}-atom FCC argon block of the {cfg['n_atoms']}-atom workload per step (same a=5.41, rc=2.5sigma, dt=0.25, "
               f"T0={args.temperature}K)") +
              f", OpenMP all-core variant of the LJVOffsetManager loop on {cores} threads, {args.steps} timed steps after {args.warmup}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "sample_atoms": n, "sample_is_whole_workload": whole,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus: int) -> dict:
    n_atoms = 4 * args.ncell ** 3 * n_gpus if not (args.strong and n_gpus > 1) else 4 * args.ncell_multi ** 3
    return {"workload": f"synthetic FCC argon {n_atoms} atoms (a=5.41), LJ rc=2.5sigma "
                        f"skin=0.3sigma, NVE dt=0.25, T0={args.temperature}K",
            "n_atoms": n_atoms, "rc": RC, "skin": SKIN, "dt": DT, "T0": args.temperature,
            "l2_policy": "working set (state + neighbour list, GBs) >> 126 MB L2; no explicit flush",
            "parallelism": "1 GPU" if n_gpus == 1 else f"spatial decomposition over {n_gpus} GPUs"}


STRONG_FILE = os.path.join(ROOT, "gpurun_out", "strong_32M_n1.json")


def strong_leg_single(args) -> dict:
    """BASELINE configs[3] at N = 1: the 32M-atom system (200^3 FCC cells) on ONE GPU, device-resident NVE steps timed with
    CUDA events on the library stream -- the denominator of the strong-scaling curve the N > 1 runs report."""
    import torch

    from pis_b200 import LennardJones, LJCudaManager
    from pis_b200.lattice import fcc_argon

    atoms = fcc_argon(args.ncell_multi, temperature=args.temperature, seed=12345)
    n = atoms.n_atoms
    mgr = LJCudaManager(skin=SKIN, device=0)
    mgr.insert((1, 1), LennardJones(0.238, SIGMA, RC, True))
    mgr.attach(atoms)
    del atoms
    mgr.compute()
    stream = torch.cuda.ExternalStream(mgr.stream_ptr)
    mgr.step_nve(DT, max(args.strong_warmup, 3))
    mgr.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    mgr.step_nve(DT, args.strong_steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.strong_steps
    st = mgr.stats()
    mgr.close()
    out = {"value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": args.strong_steps, "warmup": max(args.strong_warmup, 3),
           "n_atoms": n, "n_gpus": 1, "device_bytes": st["device_bytes"],
           "workload": f"synthetic FCC argon {n} atoms ({args.ncell_multi}^3 cells), same potential / dt / T0: BASELINE configs[3] on 1 GPU"}
    try:
        os.makedirs(os.path.dirname(STRONG_FILE), exist_ok=True)
        with open(STRONG_FILE, "w") as f:
            json.dump(out, f)
    except OSError:
        pass
    return out


def run_single(args):
    import torch

    from pis_b200 import LennardJones, LJCudaManager, capi
    from pis_b200.lattice import fcc_argon

    if capi.load().pisb_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.init()
    t_wall0 = time.perf_counter()
    ncell = args.ncell
    atoms = fcc_argon(ncell, temperature=args.temperature, seed=12345, pinned=True)
    n = atoms.n_atoms
    mgr = LJCudaManager(skin=SKIN, device=0)
    mgr.insert((1, 1), LennardJones(0.238, SIGMA, RC, True))
    options = {}
    for kv in args.option:          # library options for A/B runs, e.g. --option fuse_vv=0
        name, _, val = kv.partition("=")
        options[name] = float(val)
        mgr.set_option(name, float(val))
    fv = options.get("force_variant", 0.0)   # mirrors quad_mode() / fused_step_possible() of pisb_sim.cu
    quad = fv == 7.0 or (fv == 0.0 and n <= 75000)
    fused = options.get("fuse_vv", 1.0) != 0.0 and (quad or fv == 3.0 or fv == 0.0)
    mgr.attach(atoms)
    mgr.compute()
    stream = torch.cuda.ExternalStream(mgr.stream_ptr)
    # ---- warm-up (untimed) ----
    if args.warmup > 0:
        mgr.step_nve(DT, args.warmup)
    mgr.synchronize()
    # ---- timed region: K steps through the default launch path (CUDA-graph replays of 8/4/2 steps, the rebuild
    #      chain inside a device-side conditional node), CUDA events on the library's stream ----
    st0 = mgr.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clk:
        torch.cuda.synchronize()
        ev0.record(stream)
        th = mgr.step_nve(DT, args.steps)
        ev1.record(stream)
        torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    st1 = mgr.stats()
    ms_per_step = ms_total / args.steps
    value = n * args.steps / (ms_total * 1e-3)
    launches = st1["n_launches"] - st0["n_launches"]
    builds = st1["n_builds"] - st0["n_builds"]
    # ---- per-kernel pass: the next K steps with a CUDA-event pair around every launch (classic launch sequence: event
    #      pairs cannot be read back per replay from inside a graph); feeds the roofline and the per-kernel table ----
    mgr.set_profiling(True)
    mgr.timings(reset=True)
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record(stream)
    mgr.step_nve(DT, args.steps)
    evp1.record(stream)
    torch.cuda.synchronize()
    ms_profiled = evp0.elapsed_time(evp1)
    tim = mgr.timings()
    mgr.set_profiling(False)
    st2 = mgr.stats()

    # ---- roofline of the dominant kernel (one launch per step: LJ force pass + velocity-Verlet kick + drift).
    #      ncu (profiles/r02_force_vv_lean_full.*) shows what binds it: the L1TEX data pipe (90.9 % of peak) -- not HBM (23 %), not
    #      the FP64 pipe (43 % after the lean loop).  The primary `roofline` is therefore the L1TEX one; the HBM view
    #      (algorithmic bytes / measured copy peak) and the FP64 view are reported beside it as `roofline_hbm` / `roofline_fp64`.
    #      Algorithmic bytes per atom: LJ force 48 + 4K for a full list of K 4-byte indices; the fused step
    #      kernel adds the integrator's streams: kick 72 (v read + write, F(t) read) and, on every launch but the last of a
    #      batch, drift 72 (x_build read, x + FP32 shadow write) -> 192 + 4K. ----
    peaks, peak_kind = measured_peaks()
    ls = mgr.list_stats()
    k_mean, k_in, words = ls["listed"] / n, ls["in_range"] / n, ls["index_words"] / n
    f_ms = tim["force"]["ms"] / max(tim["force"]["launches"], 1)
    force_launches = max(tim["force"]["launches"], 1)
    drift_frac = max(force_launches - 1, 0) / force_launches      # the profiled pass is one batch: its last launch does not drift
    per_atom = 48.0 + 4.0 * k_mean + ((72.0 + 72.0 * drift_frac) if fused else 0.0)
    per_atom_stored = per_atom - 4.0 * k_mean + 4.0 * words
    bytes_per_launch = per_atom * n
    kernel_name = (("k_force_q" if quad else "k_force_vv") + "<fused>") if fused else ("k_force_q" if quad else "k_force_v3")
    achieved = bytes_per_launch / (f_ms * 1e-3) / 1e9
    peak = float(peaks.get("hbm_gbs", 6650.0))
    tj, traffic, fp64_pct, traffic_src = force_calibration(kernel_name)
    timing_note = ("per-launch CUDA events over the K steps that follow the timed region (same state, same kernels; "
                   "ms_per_step_profiled beside ms_per_step)")
    roofline_fp64 = fp64_roofline(kernel_name, k_mean, k_in, n, f_ms, fused)
    roofline_fp64.update({"timing": timing_note, "fp64_pipe_active_pct_ncu": fp64_pct})
    roofline_hbm = {"kernel": kernel_name, "bound": "hbm", "timing": timing_note,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_kind,
                    "algorithmic_bytes_per_atom": per_atom, "index_words_per_atom": words,
                    "bytes_per_atom_with_the_stored_list": per_atom_stored,
                    "frac_with_the_stored_list": per_atom_stored * n / (f_ms * 1e-3) / 1e9 / peak,
                    "mean_neighbours": k_mean, "ms_per_launch": f_ms,
                    "note": "not the binding resource: the force pass is bound by the L1TEX data pipe (roofline_l1tex; ncu: "
                            "profiles/); frac = algorithmic HBM bytes (SURVEY 8d, full list) / measured copy peak"}
    # the primary `roofline` is the resource that binds the kernel: the L1TEX data pipe
    roofline = l1tex_roofline(kernel_name, k_mean, n, f_ms, clk.summary().get("sm_mhz"), tj)
    roofline.update({"timing": timing_note, "share_of_step": tim["force"]["ms"] / ms_profiled, "traffic": traffic,
                     "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "hbm_frac": achieved / peak, "fp64_frac": roofline_fp64["frac"]})
    kernel_ms = {k: round(v["ms"] / args.steps, 5) for k, v in tim.items() if v["launches"]}
    # the streaming kernel next to it: k_vv<kick,drift> moves 200 B/atom per step; with the fused
    # step kernel only the drift that opens a batch is left (x 32 + v 24 + F 24 + x_build 24 read, x 32 + shadow 16 written)
    vv_ms = tim["integrate"]["ms"] / max(tim["integrate"]["launches"], 1)
    vv_bytes = 152.0 if fused else 200.0
    roofline_integrate = {"kernel": "k_vv<drift> (opens a batch)" if fused else "k_vv<kick,drift>", "bound": "hbm",
                          "achieved": vv_bytes * n / (vv_ms * 1e-3) / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": vv_bytes * n / (vv_ms * 1e-3) / 1e9 / peak, "ms_per_launch": vv_ms,
                          "launches": tim["integrate"]["launches"], "algorithmic_bytes_per_atom": vv_bytes}

    roofline_build = build_roofline(tim, int(st2["n_builds"] - st1["n_builds"]), n, k_mean)

    # ---- the link the end-to-end call lives on: pinned 3N-double copies each way (torch plumbing, no product code) ----
    pcie = measure_pcie(atoms.forces)

    # ---- end-to-end: the reference-facing trait call with HOST buffers, every step ----
    e2e_steps = max(3, min(args.e2e_steps, args.steps))
    mgr.download(atoms)
    mgr.verlet_step_nve(atoms, DT)  # warm the staging buffers
    mgr.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mgr.verlet_step_nve(atoms, DT)
    mgr.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": n * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 72 * n, "d2h_bytes_per_step": 72 * n + 32,
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "call": "LJCudaManager.verlet_step_nve(atoms, dt) == pisb_verlet_step_nve_host: pinned host x,v,F up in chunks, "
                   "drift per chunk, x(t+dt) down under the upload, force + kick, v,F + PE down in chunks",
           "pcie": pcie,
           # x, v, F up, then (after the force pass) v, F down; x(t+dt) can hide under the upload.  The force pass (and a
           # rebuild, when one is due) sits between the two transfers and overlaps neither: floor = link time + step time
           "pcie_floor_ms": 1e3 * (72.0 * n / (pcie["h2d_GBps"] * 1e9) + 48.0 * n / (pcie["d2h_GBps"] * 1e9)),
           "pcie_plus_step_floor_ms": 1e3 * (72.0 * n / (pcie["h2d_GBps"] * 1e9) + 48.0 * n / (pcie["d2h_GBps"] * 1e9)) + ms_per_step}

    # ---- the same loop the C++ host's Simulation::run drives (device-resident; reported beside the strict e2e) ----
    mgr.attach(atoms)
    mgr.compute()
    mgr.step_nve(DT, 10)
    mgr.download_begin(atoms, positions=True, velocities=False, forces=False)   # untimed: allocates the snapshot buffer
    mgr.download_end()
    mgr.synchronize()
    t0 = time.perf_counter()
    d2h_res = 0
    res_steps = 0
    res_total = max(e2e_steps, 60)               # >= 6 dump intervals: a 20-step window holds 3 or 4 rebuilds (+-8 % on its own)
    while res_steps < res_total:                 # Simulation::run of pis_host.cpp: one batch per dump interval
        chunk = min(10, res_total - res_steps)   # dump cadence of example/input.pis
        mgr.step_nve(DT, chunk)                  # thermo records (32 B per step) come back with the batch
        mgr.download_end()                       # the previous dump frame travelled while this batch ran
        d2h_res += 32 * chunk
        res_steps += chunk
        if res_steps % 10 == 0:
            mgr.download_begin(atoms, positions=True, velocities=False, forces=False)
            d2h_res += 24 * n
    mgr.download_end()
    mgr.synchronize()
    res_s = time.perf_counter() - t0
    e2e_resident = {"value": n * res_total / res_s, "unit": UNIT, "ms_per_step": 1e3 * res_s / res_total, "steps": res_total,
                    "d2h_bytes_per_step": d2h_res // res_total, "h2d_bytes_per_step": 0,
                    "call": "Simulation::run loop of the C++ host: state uploaded once, pisb_step_nve(dt, steps to the next dump) "
                            "returning one thermo record per step, positions of every 10th step snapshotted on the device and "
                            "copied back (pisb_download_begin/_end) while the next batch runs"}

    # ---- CPU baseline: oracle port on a bounded sample ----
    cpu = None
    if not args.no_cpu_baseline:
        cvalue, cores, per = time_cpu_path(args.cpu_ncell, args.temperature, args.cpu_steps, 1, args.cpu_threads)
        svalue, _, sper = time_cpu_path(args.ref_ncell_serial, args.temperature, 10, 1, 1)
        cpu = {"value": cvalue, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{4 * args.cpu_ncell ** 3}-atom FCC argon block of the workload (same lattice, cutoff, dt, T0), {args.cpu_steps} steps, "
                         f"OpenMP all-core variant on {cores} threads ({sum(per):.1f} s)",
               "serial_value": svalue, "serial_sample": f"{4 * args.ref_ncell_serial ** 3} atoms, 10 steps, 1 thread ({sum(sper):.1f} s) "
                                                        f"(what the reference actually executes)"}

    h = th["pe"] + th["ke"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, 1), "clocks": clk.summary(), "e2e": e2e,
        "e2e_resident": e2e_resident,
        "gpu_launches": int(launches), "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_fp64": roofline_fp64,
        "roofline_integrate": roofline_integrate, "roofline_build": roofline_build, "cpu_baseline": cpu,
        "kernel_ms_per_step": kernel_ms, "ms_per_step_profiled": ms_profiled / args.steps,
        "list_builds_in_timed_region": int(builds), "list_builds_in_profiled_pass": int(st2["n_builds"] - st1["n_builds"]),
        "launch_path": "CUDA graphs (8/4/2 steps per replay, conditional rebuild node); gpu_launches counts executed kernel nodes",
        "options": options,
        "energy_drift_rel": float(np.abs(h - h[0]).max() / abs(h[0])),
        "stats": st1, "wall_s": time.perf_counter() - t_wall0,
    }
    mgr.close()
    del atoms
    if not args.no_strong:
        try:
            line["strong_32M"] = strong_leg_single(args)
        except Exception as e:  # the headline line must not be lost to the extra leg
            line["strong_32M"] = {"error": repr(e)[:300]}
        line["wall_s"] = time.perf_counter() - t_wall0
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncell", type=int, default=100, help="FCC cells per edge at N=1 (100 -> 4M atoms)")
    ap.add_argument("--ncell-multi", type=int, default=200, help="FCC cells per edge at N>1 (200 -> 32M atoms)")
    ap.add_argument("--temperature", type=float, default=43.0, help="initial temperature (K); 43 K exercises rebuilds")
    ap.add_argument("--ref-ncell", type=int, default=0, help="--impl reference: FCC cells per edge of the CPU sample; 0 = the real per-GPU block (--ncell)")
    ap.add_argument("--ref-budget-s", type=float, default=300.0, help="--impl reference: fall back to a 256k-atom block if K + W real-size steps would take longer")
    ap.add_argument("--cpu-ncell", type=int, default=40, help="cpu_baseline leg of the GPU arm: CPU sample size (40 -> 256k atoms)")
    ap.add_argument("--ref-ncell-serial", type=int, default=20)
    ap.add_argument("--cpu-steps", type=int, default=40, help="CPU-baseline sample: steps of the ref-ncell block (~10 s on 16 cores)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--option", action="append", default=[], help="library option name=value (pisb_set_option), repeatable")
    ap.add_argument("--strong", action="store_true", help="N>1: make the 32M-atom system the PRIMARY leg instead of 4M atoms per GPU")
    ap.add_argument("--no-strong", action="store_true", help="skip the second, separately timed 32M-atom leg (`strong_32M`)")
    ap.add_argument("--strong-steps", type=int, default=10)
    ap.add_argument("--strong-warmup", type=int, default=3)
    ap.add_argument("--no-multi-parity", action="store_true", help="N>1: skip the multi-GPU == single-GPU check that precedes the timed region")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        from pis_b200 import multigpu_bench

        return multigpu_bench.run(args)
    return run_single(args)


if __name__ == "__main__":
    main()
