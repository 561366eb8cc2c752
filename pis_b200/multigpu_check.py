"""P-GPU run == 1-GPU run on the same input, as a function (used by tools/multi_check.py, by the -m gpu suite through that
script, and by bench.py --gpus N before its timed region so the driver's scaling record carries the check).

Every rank builds its brick of an FCC-argon system, runs compute + NVE steps spatially decomposed, and rank 0 additionally
runs the same global system on a single-GPU handle; compared per global id: neighbour sets (exact), forces (1e-10), thermo
traces (1e-9)."""
from __future__ import annotations

import numpy as np


def run_check(rank: int, world: int, local: int, ncell: int = 16, steps: int = 60, T0: float = 60.0, halo_mode: int = 0,
              force_variant: int = 0, fuse_vv: int | None = None, options: dict | None = None) -> dict | None:
    import torch.distributed as dist

    from . import Atoms, LennardJones, LJCudaManager, SimulationBox
    from .decomposition import create_velocities_distributed, fcc_brick, grid_for
    from .distributed import DistributedLJ, allreduce_sum_host, gather_by_gid
    from .lattice import ARGON, fcc_argon

    grid = grid_for(world)
    a = ARGON["a"]
    L = ncell * a
    rc, skin = 2.5 * ARGON["sigma"], 0.3 * ARGON["sigma"]
    n_global = 4 * ncell ** 3
    box = SimulationBox.from_lammps_data(0, L, 0, L, 0, L)

    pos, gid = fcc_brick(ncell, rank, grid)
    m = np.full(len(gid), ARGON["mass"])
    vel = create_velocities_distributed(gid, m, T0, 777, n_global, allreduce_sum_host)
    atoms = Atoms(np.ones(len(gid), dtype=np.int32), [ARGON["mass"]], pos, box, velocities=vel)
    mgr = DistributedLJ(skin=skin, local_device=local, rank=rank, world=world, grid=grid)
    mgr.insert((1, 1), LennardJones(ARGON["epsilon"], ARGON["sigma"], rc, True))
    opts = dict(options or {})
    opts["halo_mode"] = halo_mode
    if force_variant:
        opts["force_variant"] = force_variant
    if fuse_vv is not None:
        opts["fuse_vv"] = fuse_vv
    for k, v in opts.items():
        mgr.set_option(k, v)
    mgr.attach_owned(atoms, gid)
    pe0 = mgr.compute()
    g0, _, _, f0 = (a_.copy() for a_ in mgr.download_owned())
    gr0, rows0 = mgr.neighbours_owned()
    th = mgr.step_nve(0.25, steps)
    g1, x1, v1, f1 = (a_.copy() for a_ in mgr.download_owned())
    # the asynchronous form of the same download must deliver the same frame -- a snapshot: the batch that follows the begin
    # (run below as the first steps of the NVT leg) must not leak into it
    ga, xa, va, fa = mgr.download_owned_begin()
    st = mgr.stats()
    (F0,) = gather_by_gid(g0, [f0], n_global)
    X1, V1, F1 = gather_by_gid(g1, [x1, v1, f1], n_global)
    # verlet_step_nvt_nhc on bricks (pisb_step_nvt_nhc, collective): Nose-Hoover chain replicated per rank, fed with the
    # all-reduced kinetic energy; 30 steps of a 60 K -> 90 K ramp continuing from the NVE state.  tau = 100 (the value of the
    # reference's example script): with tau = 25 at dt = 0.25 the reference's chain (xi assigned, not incremented) overshoots
    # so far that the CPU restatement's own kinetic energy swings between 0.0 and 7e7 within 30 steps -- nothing to compare.
    chain = mgr.nhc_new(T0, 1.5 * T0, 100.0)
    th_nvt, en_nvt = mgr.step_nvt_nhc(0.25, 30, chain, 0, 30)
    mgr.download_end()
    oa, o1 = np.argsort(ga), np.argsort(g1)
    async_ok = bool(np.array_equal(ga[oa], g1[o1]) and np.array_equal(xa[oa], x1[o1]) and np.array_equal(va[oa], v1[o1])
                    and np.array_equal(fa[oa], f1[o1]))
    g3, x3, v3, _ = (a_.copy() for a_ in mgr.download_owned(forces=False))
    X3, V3 = gather_by_gid(g3, [x3, v3], n_global)
    # `velocity all create` on the device (pisb_start_velocities, collective): the bricks must get, per global id, the velocities
    # one GPU generates (id-keyed generator, all-reduced drift / kinetic-energy sums)
    mgr.start_velocities(35.0, 4242)
    g2, _, v2, _ = (a_.copy() for a_ in mgr.download_owned(positions=False, forces=False))
    (V2,) = gather_by_gid(g2, [v2], n_global)
    pieces = [None] * world
    dist.all_gather_object(pieces, (gr0, rows0, st, async_ok))
    out = None
    if rank == 0:
        # the same system on ONE GPU
        ref = fcc_argon(ncell, temperature=0.0)
        mref = np.full(n_global, ARGON["mass"])
        ref.velocities[...] = create_velocities_distributed(np.arange(n_global), mref, T0, 777, n_global)
        single = LJCudaManager(skin=skin, device=local)
        single.insert((1, 1), LennardJones(ARGON["epsilon"], ARGON["sigma"], rc, True))
        for k, v in opts.items():
            if k != "halo_mode":
                single.set_option(k, v)
        single.attach(ref)
        pe0_ref = single.compute()
        single.download(ref, positions=False, velocities=False)
        f0_ref = ref.forces.copy()
        rows_ref = single.neighbours(n_global)
        th_ref = single.step_nve(0.25, steps)
        single.download(ref)
        x_end, v_end, f_end = ref.positions.copy(), ref.velocities.copy(), ref.forces.copy()
        chain_ref = single.nhc_new(T0, 1.5 * T0, 100.0)
        th_nvt_ref, en_nvt_ref = single.step_nvt_nhc(0.25, 30, chain_ref, 0, 30)
        single.download(ref)
        nvt = {
            "pe_trace_rel": float(np.max(np.abs(th_nvt["pe"] - th_nvt_ref["pe"]) / np.abs(th_nvt_ref["pe"]))),
            "ke_trace_rel": float(np.max(np.abs(th_nvt["ke"] - th_nvt_ref["ke"]) / np.abs(th_nvt_ref["ke"]))),
            "thermostat_energy_rel": float(np.max(np.abs(en_nvt - en_nvt_ref) / np.maximum(np.abs(en_nvt_ref), 1e-3))),
            "xi_rel": float(max(abs(a_ - b_) / max(abs(b_), 1e-12) for a_, b_ in zip(chain.xi, chain_ref.xi))),
            "pos_max_abs": float(np.abs(X3 - ref.positions).max()), "vel_max_abs": float(np.abs(V3 - ref.velocities).max()),
        }
        nvt["ok"] = bool(nvt["pe_trace_rel"] < 1e-9 and nvt["ke_trace_rel"] < 1e-9 and nvt["thermostat_energy_rel"] < 1e-9
                         and nvt["xi_rel"] < 1e-9 and nvt["pos_max_abs"] < 1e-9)
        single.start_velocities(35.0, 4242)
        single.download(ref, positions=False, forces=False)
        v_init_err = float(np.abs(V2 - ref.velocities).max() / np.abs(ref.velocities).max())
        ref.positions[...], ref.velocities[...], ref.forces[...] = x_end, v_end, f_end
        mism = 0
        for g, rows, _, _ in pieces:
            for k, gi in enumerate(g):
                if not np.array_equal(rows[k], rows_ref[gi]):
                    mism += 1
        mag = np.linalg.norm(ref.forces, axis=1)
        den = np.maximum(mag, 1e-3 * np.sqrt((mag ** 2).mean()))
        out = {
            "world": world, "grid": list(grid), "n_global": n_global, "steps": steps, "halo_mode": halo_mode, "force_variant": force_variant,
            "rows_mismatching": mism, "neighbour_rows_mismatching": mism,
            "force0_max_abs": float(np.abs(F0 - f0_ref).max()),
            "force_rel": float((np.linalg.norm(F1 - ref.forces, axis=1) / den).max()),
            "pe0_rel": abs(pe0 - pe0_ref) / abs(pe0_ref),
            "pe_trace_rel": float(np.max(np.abs(th["pe"] - th_ref["pe"]) / np.abs(th_ref["pe"]))),
            "ke_trace_rel": float(np.max(np.abs(th["ke"] - th_ref["ke"]) / np.abs(th_ref["ke"]))),
            "virial_ref_trace_rel": float(np.max(np.abs(th["virial_ref"] - th_ref["virial_ref"]) / np.maximum(np.abs(th_ref["virial_ref"]), 1.0))),
            "pos_max_abs": float(np.abs(X1 - ref.positions).max()),
            "vel_max_abs": float(np.abs(V1 - ref.velocities).max()),
            "start_velocities_rel": v_init_err, "nvt": nvt, "async_download_is_a_snapshot": all(p[3] for p in pieces),
            "builds_multi": [p[2]["n_builds"] for p in pieces], "builds_single": single.stats()["n_builds"],
            "owned": [p[2]["n_atoms"] for p in pieces], "ghost": [p[2]["n_ghost"] for p in pieces],
        }
        ok = (mism == 0 and out["force_rel"] < 1e-10 and out["force0_max_abs"] < 1e-12 and out["pe0_rel"] < 1e-9 and out["pe_trace_rel"] < 1e-9
              and out["ke_trace_rel"] < 1e-9 and v_init_err < 1e-12 and nvt["ok"] and out["async_download_is_a_snapshot"])
        out["ok"] = bool(ok)
        single.close()
    dist.barrier()
    mgr.close()
    return out
