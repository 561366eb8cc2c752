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
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

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
                "samples": len(self.rows), "reasons": reasons}


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


# FP64 pipe yardstick (the north_star's alternative to the HBM fraction).  Issue slots of the force loop per listed pair,
# counted in the SASS of k_force_v3 / k_force_vv: 14 distance + 5 reciprocal + 11 LJ + 5 accumulate;
# the fused kernel's epilogue adds ~110 per atom (6 IEEE divisions, kick, drift, wrap, skin trigger).  Peak: the measured
# DFMA rate of tools/microbench.cu, 57.2 per clock and SM (profiles/r01_microbench_pipes.txt), 148 SMs at 1965 MHz.
FP64_SLOTS_PER_PAIR = 35.0
FP64_SLOTS_EPILOGUE = 110.0
FP64_PEAK_SLOTS_PER_S = 57.16 * 148 * 1.965e9


def fp64_roofline(k_mean: float, n_atoms: int, ms_per_launch: float, fused: bool) -> dict:
    slots = FP64_SLOTS_PER_PAIR * k_mean + (FP64_SLOTS_EPILOGUE if fused else 0.0)
    achieved = slots * n_atoms / (ms_per_launch * 1e-3)
    return {"bound": "fp64 pipe", "slots_per_atom": slots, "achieved": achieved / 1e12, "peak": FP64_PEAK_SLOTS_PER_S / 1e12,
            "unit": "T FP64 instr/s", "frac": achieved / FP64_PEAK_SLOTS_PER_S,
            "peak_source": "measured DFMA rate (tools/microbench.cu), 57.2 per clock and SM"}


def argon_oracle(atoms):
    from oracle.pis_oracle import Oracle

    orc = Oracle(atoms.sim_box.h_colmajor(), masses=atoms.masses)
    orc.insert(1, 1, 0.238, SIGMA, RC)
    return orc


def time_cpu_path(ncell: int, temperature: float, steps: int, warmup: int, threads: int, seed: int = 12345):
    """The reference's CPU path (oracle port, OpenMP all-core variant of the LJVOffsetManager loop) on an
    FCC-argon system of ncell^3*4 atoms: returns (atom-steps/s, cores, seconds per step list)."""
    from pis_b200.lattice import fcc_argon

    atoms = fcc_argon(ncell, temperature=temperature, seed=seed)
    orc = argon_oracle(atoms)
    cores = threads if threads > 0 else orc.max_threads()
    mode = "omp" if cores > 1 else "serial"
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    f = np.zeros_like(x)
    orc.compute_potential(x, atoms.type_ids, forces=f, mode=mode, threads=cores)
    for _ in range(warmup):
        orc.verlet_step_nve(x, v, f, atoms.type_ids, DT, mode=mode, threads=cores)
    per = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.verlet_step_nve(x, v, f, atoms.type_ids, DT, mode=mode, threads=cores)
        per.append(time.perf_counter() - t0)
    tot = sum(per)
    return atoms.n_atoms * steps / tot, cores, per


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust reference cannot be
    built here (no rustc), so this is the oracle port, with all host threads, each step a bounded sample
    (a 256k-atom FCC-argon block: same lattice, density, cutoff and dt as the 4M workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncell = args.ref_ncell
    t0 = time.perf_counter()
    value, cores, per = time_cpu_path(ncell, args.temperature, args.steps, args.warmup, args.cpu_threads)
    n = 4 * ncell ** 3
    ms = 1e3 * sum(per) / len(per)
    sample = (f"{n}-atom FCC argon block per step (same a=5.41, rc=2.5sigma, dt=0.25, T0={args.temperature}K as the "
              f"{4 * args.ncell ** 3}-atom workload), OpenMP all-core variant of the LJVOffsetManager loop, "
              f"{args.steps} timed steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
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
    fv = options.get("force_variant", 0.0)   # mirrors fused_step_possible() of pisb_sim.cu
    fused = options.get("fuse_vv", 1.0) != 0.0 and (fv == 3.0 or (fv == 0.0 and n > 75000))
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

    # ---- roofline of the dominant kernel.  Algorithmic bytes per atom: LJ force 48 + 4K; the fused step
    #      kernel k_force_vv adds the integrator's streams: kick 72 (v read + write, F(t) read) and, on every launch but
    #      the last of a batch, drift 72 (x_build read, x + FP32 shadow write) -> 192 + 4K ----
    peaks, peak_kind = measured_peaks()
    nn = np.zeros(n, dtype=np.int32)
    capi.check(mgr._h, capi.load().pisb_neighbours(mgr._h, capi._ptr(nn), None, 0))
    k_mean = float(nn.mean())
    f_ms = tim["force"]["ms"] / max(tim["force"]["launches"], 1)
    force_launches = max(tim["force"]["launches"], 1)
    drift_frac = max(force_launches - 1, 0) / force_launches      # the profiled pass is one batch: its last launch does not drift
    per_atom = 48.0 + 4.0 * k_mean + ((72.0 + 72.0 * drift_frac) if fused else 0.0)
    bytes_per_launch = per_atom * n
    kernel_name = "k_force_vv" if fused else "k_force_v3"
    achieved = bytes_per_launch / (f_ms * 1e-3) / 1e9
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, fp64_pct = None, None
    tp = os.path.join(ROOT, "profiles", "force_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                tj = json.load(f)
            if str(tj.get("kernel", "")).startswith(kernel_name):    # a capture of another kernel says nothing about this one
                traffic, fp64_pct = tj.get("dram_bytes_per_launch"), tj.get("fp64_pipe_active_pct")
        except Exception:
            traffic = None
    roofline = {"kernel": kernel_name, "bound": "hbm", "timing": "per-launch CUDA events over the K steps that follow the timed "
                "region (same state, same kernels; ms_per_step_profiled beside ms_per_step)",
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "fp64_pipe_active_pct_ncu": fp64_pct, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                "algorithmic_bytes_per_atom": per_atom, "mean_neighbours": k_mean,
                "ms_per_launch": f_ms, "share_of_step": tim["force"]["ms"] / ms_profiled,
                "note": ("one launch per step: LJ force pass with the velocity-Verlet kick + drift in its epilogue; " if fused else "") +
                        "the force pass is bound by the FP64 pipe and L1TEX gather throughput, not by HBM (ncu: profiles/); "
                        "frac is algorithmic HBM bytes / measured copy peak; traffic = ncu dram bytes per launch"}
    kernel_ms = {k: round(v["ms"] / args.steps, 5) for k, v in tim.items() if v["launches"]}
    # the streaming kernel next to it: k_vv<kick,drift> moves 200 B/atom per step; with the fused
    # step kernel only the drift that opens a batch is left (x 32 + v 24 + F 24 + x_build 24 read, x 32 + shadow 16 written)
    vv_ms = tim["integrate"]["ms"] / max(tim["integrate"]["launches"], 1)
    vv_bytes = 152.0 if fused else 200.0
    roofline_integrate = {"kernel": "k_vv<drift> (opens a batch)" if fused else "k_vv<kick,drift>", "bound": "hbm",
                          "achieved": vv_bytes * n / (vv_ms * 1e-3) / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": vv_bytes * n / (vv_ms * 1e-3) / 1e9 / peak, "ms_per_launch": vv_ms,
                          "launches": tim["integrate"]["launches"], "algorithmic_bytes_per_atom": vv_bytes}

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
    while res_steps < e2e_steps:                 # Simulation::run of pis_host.cpp: one batch per dump interval
        chunk = min(10, e2e_steps - res_steps)   # dump cadence of example/input.pis
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
    e2e_resident = {"value": n * e2e_steps / res_s, "unit": UNIT, "ms_per_step": 1e3 * res_s / e2e_steps,
                    "d2h_bytes_per_step": d2h_res // e2e_steps, "h2d_bytes_per_step": 0,
                    "call": "Simulation::run loop of the C++ host: state uploaded once, pisb_step_nve(dt, steps to the next dump) "
                            "returning one thermo record per step, positions of every 10th step snapshotted on the device and "
                            "copied back (pisb_download_begin/_end) while the next batch runs"}

    # ---- CPU baseline: oracle port on a bounded sample ----
    cpu = None
    if not args.no_cpu_baseline:
        cvalue, cores, per = time_cpu_path(args.ref_ncell, args.temperature, args.cpu_steps, 1, args.cpu_threads)
        svalue, _, sper = time_cpu_path(args.ref_ncell_serial, args.temperature, 10, 1, 1)
        cpu = {"value": cvalue, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{4 * args.ref_ncell ** 3}-atom FCC argon block, {args.cpu_steps} steps, OpenMP all-core variant "
                         f"({sum(per):.1f} s)",
               "serial_value": svalue, "serial_sample": f"{4 * args.ref_ncell_serial ** 3} atoms, 10 steps, 1 thread ({sum(sper):.1f} s) "
                                                        f"(what the reference actually executes)"}

    h = th["pe"] + th["ke"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, 1), "clocks": clk.summary(), "e2e": e2e,
        "e2e_resident": e2e_resident,
        "gpu_launches": int(launches), "roofline": roofline, "roofline_fp64": fp64_roofline(k_mean, n, f_ms, fused),
        "roofline_integrate": roofline_integrate, "cpu_baseline": cpu,
        "kernel_ms_per_step": kernel_ms, "ms_per_step_profiled": ms_profiled / args.steps,
        "list_builds_in_timed_region": int(builds), "list_builds_in_profiled_pass": int(st2["n_builds"] - st1["n_builds"]),
        "launch_path": "CUDA graphs (8/4/2 steps per replay, conditional rebuild node); gpu_launches counts executed kernel nodes",
        "options": options,
        "energy_drift_rel": float(np.abs(h - h[0]).max() / abs(h[0])),
        "stats": st1, "wall_s": time.perf_counter() - t_wall0,
    }
    print(json.dumps(line), flush=True)
    mgr.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncell", type=int, default=100, help="FCC cells per edge at N=1 (100 -> 4M atoms)")
    ap.add_argument("--ncell-multi", type=int, default=200, help="FCC cells per edge at N>1 (200 -> 32M atoms)")
    ap.add_argument("--temperature", type=float, default=43.0, help="initial temperature (K); 43 K exercises rebuilds")
    ap.add_argument("--ref-ncell", type=int, default=40, help="CPU sample size (40 -> 256k atoms)")
    ap.add_argument("--ref-ncell-serial", type=int, default=20)
    ap.add_argument("--cpu-steps", type=int, default=40, help="CPU-baseline sample: steps of the ref-ncell block (~10 s on 16 cores)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--option", action="append", default=[], help="library option name=value (pisb_set_option), repeatable")
    ap.add_argument("--strong", action="store_true", help="N>1: run the 32M-atom system instead of 4M atoms per GPU")
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
