"""bench.py leg for N > 1 GPUs: spatial decomposition, one rank per GPU (launched by torchrun).

One run, three parts, every timed region on the device (CUDA events on the library stream, max over ranks), profiling OFF:
  0. `multi_parity` : the N-GPU == 1-GPU check of pis_b200/multigpu_check.py (neighbour rows per global id, forces, traces)
                      on a small system stepped by the same kernels as the big bricks -- BEFORE anything is timed; a failure
                      makes the process exit non-zero.
  1. PRIMARY (weak): every rank owns one 4M-atom brick (ncell^3 FCC cells), so N = 8 is BASELINE configs[3] (32M atoms,
                      2x2x2 bricks) and N = 1 is configs[2]; `--strong` makes the 32M-atom system the primary leg instead.
  2. `strong_32M`   : the 32M-atom system on these N GPUs, separately timed (BASELINE configs[3]'s strong-scaling curve);
                      with the N = 1 value of the same box (gpurun_out/strong_32M_n1.json, written by the N = 1 run) it also
                      reports v_N / (N v_1).
The per-kernel classes come from a FOLLOWING pass with per-launch events, as at N = 1."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np


def _build(args, nc, rank, world, local, grid, seed=12345):
    from .atoms import Atoms
    from .decomposition import create_velocities_distributed, fcc_brick
    from .distributed import DistributedLJ, allreduce_sum_host
    from .lattice import ARGON
    from .potentials import LennardJones
    from .simulation_box import SimulationBox
    import bench as B

    a = ARGON["a"]
    n_global = 4 * nc[0] * nc[1] * nc[2]
    box = SimulationBox.from_lammps_data(0, nc[0] * a, 0, nc[1] * a, 0, nc[2] * a)
    pos, gid = fcc_brick(nc, rank, grid)
    m = np.full(len(gid), ARGON["mass"])
    vel = create_velocities_distributed(gid, m, args.temperature, seed, n_global, allreduce_sum_host)
    atoms = Atoms(np.ones(len(gid), dtype=np.int32), [ARGON["mass"]], pos, box, velocities=vel, pinned=True)
    del pos, vel
    mgr = DistributedLJ(skin=B.SKIN, local_device=local, rank=rank, world=world, grid=grid)
    mgr.insert((1, 1), LennardJones(ARGON["epsilon"], ARGON["sigma"], B.RC, True))
    for kv in getattr(args, "option", []):      # library options for A/B runs, e.g. --option fuse_vv=0
        name, _, val = kv.partition("=")
        mgr.set_option(name, float(val))
    mgr.attach_owned(atoms, gid)
    return mgr, atoms, n_global


def _timed(mgr, stream, steps, dist, torch):
    """steps NVE steps between barriers; returns (ms over all ranks' max, thermo)."""
    import bench as B

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    th = mgr.step_nve(B.DT, steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), th


def run(args):
    import torch
    import torch.distributed as dist

    from . import capi
    from .decomposition import grid_for
    from .distributed import init_process_group
    from .multigpu_check import run_check
    import bench as B

    rank, world = init_process_group()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    t_wall0 = time.perf_counter()
    grid = grid_for(world)

    # ---- 0. multi-GPU == single-GPU, on the driver's record ----
    multi_parity = None
    if not args.no_multi_parity:
        # 24^3 cells = 55 296 atoms at 60 K, 40 steps (rebuilds + migration); force_variant 3 = the thread-per-atom fused step
        # kernel (k_force_vv<.., BRICK>) the 4M-atom bricks use, forced at this brick size
        multi_parity = run_check(rank, world, local, ncell=24, steps=40, T0=60.0, halo_mode=0, force_variant=3)
        flag = torch.tensor([1 if (rank != 0 or multi_parity["ok"]) else 0], dtype=torch.int64)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if rank == 0:
                print(json.dumps({"metric": B.METRIC, "n_gpus": world, "error": "multi-GPU run differs from the single-GPU run",
                                  "multi_parity": multi_parity}), flush=True)
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(3)

    # ---- 1. primary leg ----
    if args.strong:
        nc, scaling = (args.ncell_multi,) * 3, "strong"
    else:
        nc, scaling = tuple(args.ncell * g for g in grid), "weak"
    mgr, atoms, n_global = _build(args, nc, rank, world, local, grid)
    h2d_total = atoms.n_atoms * (72 + 8)
    mgr.compute()
    stream = torch.cuda.ExternalStream(mgr.stream_ptr)
    if args.warmup > 0:
        mgr.step_nve(B.DT, args.warmup)
    mgr.synchronize()
    st0 = mgr.stats()
    clk = B.ClockSampler(local)
    if rank == 0:
        clk.__enter__()
    ms_total, th = _timed(mgr, stream, args.steps, dist, torch)
    if rank == 0:
        clk.__exit__()
    st1 = mgr.stats()
    value = n_global * args.steps / (ms_total * 1e-3)
    # per-kernel classes: the NEXT K steps with an event pair around every launch (as bench.py does at N = 1)
    mgr.set_profiling(True)
    mgr.timings(reset=True)
    ms_prof, _ = _timed(mgr, stream, args.steps, dist, torch)
    tim = mgr.timings()
    mgr.set_profiling(False)

    # ---- end to end: the Simulation::run loop of the host -- one batch per dump interval, thermo records back with the
    #      batch, the owned POSITIONS (+ global ids: what DumpTraj::write_step needs) back on dump steps ----
    e2e_steps = max(60, args.e2e_steps)   # >= 6 dump intervals: the frame of the LAST dump is waited for in the open (one per run)
    mgr.download_owned_begin(velocities=False, forces=False)  # untimed: allocates the pinned destination and snapshot buffers
    mgr.download_end()
    mgr.step_nve(B.DT, 10)
    st_e2e0 = mgr.stats()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d2h = 0
    s = 0
    while s < e2e_steps:
        chunk = min(10, e2e_steps - s)                       # dump cadence of example/input.pis
        mgr.step_nve(B.DT, chunk)
        mgr.download_end()                                   # the previous dump frame travelled while this batch ran
        d2h += 32 * chunk
        s += chunk
        if s % 10 == 0:
            g_, x_, v_, f_ = mgr.download_owned_begin(velocities=False, forces=False)
            d2h += x_.nbytes + g_.nbytes
    mgr.download_end()
    mgr.synchronize()
    dist.barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    st_e2e1 = mgr.stats()

    ls = mgr.list_stats()            # this rank's list: listed pairs, pairs in range, index words stored
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "stats": st1, "kernel_ms": {k: v["ms"] for k, v in tim.items() if v["launches"]}})
    n_own_now = int(mgr.stats()["n_atoms"])
    mgr.close()
    del atoms

    # ---- 2. the 32M-atom strong-scaling leg ----
    strong = None
    if not args.no_strong and nc == (args.ncell_multi,) * 3:
        # the primary leg IS the 32M-atom system (N = 8 weak, or --strong): same numbers, no second run
        strong = {"value": value, "unit": B.UNIT, "ms_per_step": ms_total / args.steps, "steps": args.steps, "warmup": args.warmup,
                  "n_atoms": n_global, "n_gpus": world, "workload": "the primary leg (32M atoms)"}
    elif not args.no_strong:
        try:
            nc_s = (args.ncell_multi,) * 3
            mgr_s, atoms_s, n_s = _build(args, nc_s, rank, world, local, grid)
            mgr_s.compute()
            stream_s = torch.cuda.ExternalStream(mgr_s.stream_ptr)
            mgr_s.step_nve(B.DT, max(args.strong_warmup, 3))
            mgr_s.synchronize()
            ms_s, _ = _timed(mgr_s, stream_s, args.strong_steps, dist, torch)
            ms_s /= args.strong_steps
            st_s = mgr_s.stats()
            mgr_s.close()
            del atoms_s
            strong = {"value": n_s / (ms_s * 1e-3), "unit": B.UNIT, "ms_per_step": ms_s, "steps": args.strong_steps,
                      "warmup": max(args.strong_warmup, 3), "n_atoms": n_s, "n_gpus": world, "owned_rank0": st_s["n_atoms"],
                      "ghost_rank0": st_s["n_ghost"],
                      "workload": f"synthetic FCC argon {n_s} atoms ({args.ncell_multi}^3 cells) on {grid[0]}x{grid[1]}x{grid[2]} bricks: BASELINE configs[3]"}
        except Exception as e:
            strong = {"error": repr(e)[:300]}
    if rank == 0 and strong and "value" in strong and os.path.exists(B.STRONG_FILE):
        try:
            with open(B.STRONG_FILE) as f:
                one = json.load(f)
            if one.get("n_atoms") == strong["n_atoms"]:
                strong["n1_value_same_box"] = one["value"]
                strong["strong_efficiency"] = strong["value"] / (world * one["value"])
        except Exception:
            pass

    if rank == 0:
        peaks, peak_kind = B.measured_peaks()
        n_own = st1["n_atoms"]
        k_mean, k_in, words = ls["listed"] / n_own_now, ls["in_range"] / n_own_now, ls["index_words"] / n_own_now
        f_ms = tim["force"]["ms"] / max(tim["force"]["launches"], 1)
        # bricks of this size step with the fused kernel (force + kick + drift): 192 + 4K algorithmic bytes per owned atom,
        # 48 + 4K for the plain force kernel (bench.py)
        fused = tim.get("integrate", {"launches": 0})["launches"] < max(args.steps // 2, 1)   # unfused: one k_vv per step
        kernel = "k_force_vv<fused,brick>" if fused else "k_force_v3"
        roofline_fp64 = B.fp64_roofline(kernel, k_mean, k_in, n_own, f_ms, fused)
        cal, traffic, _, traffic_src = B.force_calibration(kernel)
        clocks = clk.summary()
        roofline = B.l1tex_roofline(kernel, k_mean, n_own, f_ms, clocks.get("sm_mhz"), cal)   # the binding resource (bench.py)
        # traffic: the static single-GPU ncu capture of the same kernel on a 4M-atom system (a brick adds ghost gathers, not streams)
        roofline.update({"share_of_step": tim["force"]["ms"] / ms_prof, "traffic": traffic if n_own >= 3000000 else None,
                         "traffic_source": (traffic_src + "; single-GPU form of the kernel, 4M atoms") if traffic_src and n_own >= 3000000 else None,
                         "note_rank": "rank 0's brick", "fp64_frac": roofline_fp64["frac"]})
        achieved = ((192.0 if fused else 48.0) + 4.0 * k_mean) * n_own / (f_ms * 1e-3) / 1e9
        peak = float(peaks.get("hbm_gbs", 6650.0))
        cfg = B.workload_config(args, world)
        cfg["parallelism"] = (f"spatial decomposition {grid[0]}x{grid[1]}x{grid[2]} bricks, per-step halo = fused pack + NVLink stores into "
                              f"CUDA-IPC peer memory (NCCL for migration / rebuild traffic), {n_global // world} atoms per GPU")
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": n_global * e2e_steps / e2e_s, "unit": B.UNIT, "h2d_bytes_per_step": h2d_total // max(args.steps + e2e_steps + args.warmup, 1),
                    "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "call": "per rank: pisb_step_nve(dt, steps to the next dump) returning one thermo record per step and "
                            "pisb_download_owned_begin / pisb_download_end (positions + global ids) every 10 steps (the example's dump cadence, the frame travels under the next batch); "
                            "state is uploaded once"},
            "gpu_launches": int(st1["n_launches"] - st0["n_launches"]),
            "multi_parity": multi_parity, "strong_32M": strong,
            "roofline": roofline, "roofline_fp64": roofline_fp64,
            "roofline_hbm": {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": peak_kind, "mean_neighbours": k_mean, "index_words_per_atom": words,
                             "ms_per_launch": f_ms, "note": "rank 0's force kernel; not the binding resource (the L1TEX data pipe is), see DESIGN.md"},
            "cpu_baseline": None,
            "ms_per_step_profiled": ms_prof / args.steps,
            "kernel_ms_per_step_rank0": {k: round(v["ms"] / args.steps, 5) for k, v in tim.items() if v["launches"]},
            "list_builds_in_timed_region": int(st1["n_builds"] - st0["n_builds"]),
            "list_builds_in_e2e_region": int(st_e2e1["n_builds"] - st_e2e0["n_builds"]),
            "per_rank": [{"rank": g["rank"], "owned": g["stats"]["n_atoms"], "ghost": g["stats"]["n_ghost"],
                          "halo_ms_per_step": round(g["kernel_ms"].get("halo", 0.0) / args.steps, 5)} for g in gathered],
            "energy_drift_rel": float(np.abs((th["pe"] + th["ke"]) - (th["pe"][0] + th["ke"][0])).max() / abs(th["pe"][0] + th["ke"][0])),
            "launch_path": "classic launches (a brick's rebuild is host-orchestrated); timed region without per-kernel events, "
                           "kernel classes from the following pass",
            "wall_s": time.perf_counter() - t_wall0,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
