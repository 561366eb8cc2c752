"""bench.py leg for N > 1 GPUs: spatial decomposition, one rank per GPU (launched by torchrun).

Weak scaling by default: every rank owns one 4M-atom brick (ncell^3 FCC cells), so N = 8 is BASELINE
config 3 (32M atoms, 2x2x2 bricks) and N = 1 is config 2; `--strong` instead runs the 32M-atom system
on any N.  Timed on the device (CUDA events on the library stream), max over ranks."""
from __future__ import annotations

import json
import os
import time

import numpy as np


def run(args):
    import torch
    import torch.distributed as dist

    from . import capi
    from .atoms import Atoms
    from .decomposition import create_velocities_distributed, fcc_brick, grid_for
    from .distributed import DistributedLJ, allreduce_sum_host, init_process_group
    from .lattice import ARGON
    from .potentials import LennardJones
    from .simulation_box import SimulationBox
    import bench as B

    rank, world = init_process_group()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    t_wall0 = time.perf_counter()
    grid = grid_for(world)
    if args.strong:
        nc = (args.ncell_multi,) * 3
        scaling = "strong"
    else:
        nc = tuple(args.ncell * g for g in grid)
        scaling = "weak"
    a = ARGON["a"]
    n_global = 4 * nc[0] * nc[1] * nc[2]
    box = SimulationBox.from_lammps_data(0, nc[0] * a, 0, nc[1] * a, 0, nc[2] * a)
    pos, gid = fcc_brick(nc, rank, grid)
    m = np.full(len(gid), ARGON["mass"])
    vel = create_velocities_distributed(gid, m, args.temperature, 12345, n_global, allreduce_sum_host)
    atoms = Atoms(np.ones(len(gid), dtype=np.int32), [ARGON["mass"]], pos, box, velocities=vel, pinned=True)
    del pos, vel
    mgr = DistributedLJ(skin=B.SKIN, local_device=local, rank=rank, world=world, grid=grid)
    mgr.insert((1, 1), LennardJones(ARGON["epsilon"], ARGON["sigma"], B.RC, True))
    for kv in getattr(args, "option", []):      # library options for A/B runs, e.g. --option fuse_vv=0
        name, _, val = kv.partition("=")
        mgr.set_option(name, float(val))
    mgr.attach_owned(atoms, gid)
    h2d_total = atoms.n_atoms * (72 + 8)
    mgr.compute()
    stream = torch.cuda.ExternalStream(mgr.stream_ptr)
    if args.warmup > 0:
        mgr.step_nve(B.DT, args.warmup)
    mgr.synchronize()
    st0 = mgr.stats()
    mgr.set_profiling(True)
    mgr.timings(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = B.ClockSampler(local)
    if rank == 0:
        clk.__enter__()
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    th = mgr.step_nve(B.DT, args.steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        clk.__exit__()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    tim = mgr.timings()
    mgr.set_profiling(False)
    st1 = mgr.stats()
    value = n_global * args.steps / (ms_total * 1e-3)

    # ---- end to end: the Simulation::run loop of the host -- one batch per dump interval, thermo records back with the
    #      batch, the owned POSITIONS (+ global ids: what DumpTraj::write_step needs) back on dump steps ----
    e2e_steps = max(10, min(args.e2e_steps, args.steps))
    mgr.download_owned(velocities=False, forces=False)  # untimed: allocates the pinned destination buffers that every later call reuses
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
        d2h += 32 * chunk
        s += chunk
        if s % 10 == 0:
            g_, x_, v_, f_ = mgr.download_owned(velocities=False, forces=False)
            d2h += x_.nbytes + g_.nbytes
    mgr.synchronize()
    dist.barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())

    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "stats": st1, "kernel_ms": {k: v["ms"] for k, v in tim.items() if v["launches"]}})
    if rank == 0:
        peaks, peak_kind = B.measured_peaks()
        n_own = st1["n_atoms"]
        nn = np.zeros(int(mgr.stats()["n_atoms"]) + 64, dtype=np.int32)  # owned count NOW (atoms migrate)
        capi.check(mgr._h, capi.load().pisb_neighbours(mgr._h, capi._ptr(nn), None, 0))
        nn_mean = float(nn[: int(mgr.stats()['n_atoms'])].mean())
        f_ms = tim["force"]["ms"] / max(tim["force"]["launches"], 1)
        # bricks of this size step with the fused kernel (force + kick + drift): 192 + 4K algorithmic bytes per owned atom,
        # 48 + 4K for the plain force kernel (bench.py)
        fused = tim.get("integrate", {"launches": 0})["launches"] < max(args.steps // 2, 1)   # unfused: one k_vv per step
        achieved = ((192.0 if fused else 48.0) + 4.0 * nn_mean) * n_own / (f_ms * 1e-3) / 1e9
        peak = float(peaks.get("hbm_gbs", 6650.0))
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic FCC argon {n_global} atoms ({nc[0]}x{nc[1]}x{nc[2]} cells, a=5.41), LJ rc=2.5sigma "
                                   f"skin=0.3sigma, NVE dt=0.25, T0={args.temperature}K",
                       "n_atoms": n_global, "rc": B.RC, "skin": B.SKIN, "dt": B.DT, "T0": args.temperature,
                       "l2_policy": "working set per GPU (state + neighbour list, GBs) >> 126 MB L2; no explicit flush",
                       "parallelism": f"spatial decomposition {grid[0]}x{grid[1]}x{grid[2]} bricks, per-step halo = fused pack + NVLink stores into "
                                      f"CUDA-IPC peer memory (NCCL for migration / rebuild traffic), "
                                      f"{n_global // world} atoms per GPU"},
            "clocks": clk.summary(),
            "e2e": {"value": n_global * e2e_steps / e2e_s, "unit": B.UNIT, "h2d_bytes_per_step": h2d_total // max(args.steps + e2e_steps + args.warmup, 1),
                    "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "call": "per rank: pisb_step_nve(dt, steps to the next dump) returning one thermo record per step and "
                            "pisb_download_owned (positions + global ids) every 10 steps (the example's dump cadence); "
                            "state is uploaded once"},
            "gpu_launches": int(st1["n_launches"] - st0["n_launches"]),
            "roofline": {"kernel": "k_force_vv" if fused else "k_force_v3", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_kind, "mean_neighbours": nn_mean,
                         "ms_per_launch": f_ms, "share_of_step": tim["force"]["ms"] / ms_total,
                         "note": "rank 0's force kernel; FP64/L1 bound, see DESIGN.md"},
            "cpu_baseline": None,
            "kernel_ms_per_step_rank0": {k: round(v["ms"] / args.steps, 5) for k, v in tim.items() if v["launches"]},
            "list_builds_in_timed_region": int(st1["n_builds"] - st0["n_builds"]),
            "list_builds_in_e2e_region": int(mgr.stats()["n_builds"] - st_e2e0["n_builds"]),
            "per_rank": [{"rank": g["rank"], "owned": g["stats"]["n_atoms"], "ghost": g["stats"]["n_ghost"],
                          "halo_ms_per_step": round(g["kernel_ms"].get("halo", 0.0) / args.steps, 5)} for g in gathered],
            "energy_drift_rel": float(np.abs((th["pe"] + th["ke"]) - (th["pe"][0] + th["ke"][0])).max() / abs(th["pe"][0] + th["ke"][0])),
            "wall_s": time.perf_counter() - t_wall0,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    mgr.close()
    dist.destroy_process_group()
