"""P-GPU run == 1-GPU run on the same input (launched under torchrun on P GPUs; also used by the
-m gpu test-suite through subprocess when >= 2 GPUs are visible).

Every rank builds its brick of an FCC-argon system, runs compute + NVE steps spatially decomposed,
and rank 0 additionally runs the same global system on a single-GPU handle; compared per global id:
neighbour sets (exact), forces (1e-10), thermo traces (1e-9).  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pis_b200 import Atoms, LennardJones, LJCudaManager, SimulationBox  # noqa: E402
from pis_b200.decomposition import create_velocities_distributed, fcc_brick, grid_for  # noqa: E402
from pis_b200.distributed import DistributedLJ, allreduce_sum_host, gather_by_gid, init_process_group  # noqa: E402
from pis_b200.lattice import ARGON, fcc_argon  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    T0 = float(sys.argv[3]) if len(sys.argv) > 3 else 60.0
    rank, world = init_process_group()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
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
    halo_mode = int(os.environ.get("PISB_HALO_MODE", "0"))  # 0 auto (peer memory), 1 NCCL send/recv, 2 peer memory or fail
    mgr.set_option("halo_mode", halo_mode)
    force_variant = int(os.environ.get("PISB_FORCE_VARIANT", "0"))  # 3: k_force_v3 / the fused k_force_vv step at any brick size
    if force_variant:
        mgr.set_option("force_variant", force_variant)
    if "PISB_FUSE_VV" in os.environ:
        mgr.set_option("fuse_vv", int(os.environ["PISB_FUSE_VV"]))
    mgr.attach_owned(atoms, gid)
    pe0 = mgr.compute()
    g0, _, _, f0 = (a_.copy() for a_ in mgr.download_owned())
    gr0, rows0 = mgr.neighbours_owned()
    th = mgr.step_nve(0.25, steps)
    g1, x1, v1, f1 = (a_.copy() for a_ in mgr.download_owned())
    st = mgr.stats()
    (F0,) = gather_by_gid(g0, [f0], n_global)
    X1, V1, F1 = gather_by_gid(g1, [x1, v1, f1], n_global)
    pieces = [None] * world
    dist.all_gather_object(pieces, (gr0, rows0, st))

    if rank == 0:
        # the same system on ONE GPU
        ref = fcc_argon(ncell, temperature=0.0)
        mref = np.full(n_global, ARGON["mass"])
        ref.velocities[...] = create_velocities_distributed(np.arange(n_global), mref, T0, 777, n_global)
        single = LJCudaManager(skin=skin, device=local)
        single.insert((1, 1), LennardJones(ARGON["epsilon"], ARGON["sigma"], rc, True))
        single.attach(ref)
        pe0_ref = single.compute()
        single.download(ref, positions=False, velocities=False)
        f0_ref = ref.forces.copy()
        rows_ref = single.neighbours(n_global)
        th_ref = single.step_nve(0.25, steps)
        single.download(ref)
        mism = 0
        for g, rows, _ in pieces:
            for k, gi in enumerate(g):
                if not np.array_equal(rows[k], rows_ref[gi]):
                    mism += 1
        mag = np.linalg.norm(ref.forces, axis=1)
        den = np.maximum(mag, 1e-3 * np.sqrt((mag ** 2).mean()))
        out = {
            "world": world, "grid": grid, "n_global": n_global, "steps": steps, "halo_mode": halo_mode, "force_variant": force_variant,
            "neighbour_rows_mismatching": mism,
            "force0_max_abs": float(np.abs(F0 - f0_ref).max()),
            "force_rel": float((np.linalg.norm(F1 - ref.forces, axis=1) / den).max()),
            "pe0_rel": abs(pe0 - pe0_ref) / abs(pe0_ref),
            "pe_trace_rel": float(np.max(np.abs(th["pe"] - th_ref["pe"]) / np.abs(th_ref["pe"]))),
            "ke_trace_rel": float(np.max(np.abs(th["ke"] - th_ref["ke"]) / np.abs(th_ref["ke"]))),
            "virial_ref_trace_rel": float(np.max(np.abs(th["virial_ref"] - th_ref["virial_ref"]) / np.maximum(np.abs(th_ref["virial_ref"]), 1.0))),
            "pos_max_abs": float(np.abs(X1 - ref.positions).max()),
            "vel_max_abs": float(np.abs(V1 - ref.velocities).max()),
            "builds_multi": [p[2]["n_builds"] for p in pieces], "builds_single": single.stats()["n_builds"],
            "owned": [p[2]["n_atoms"] for p in pieces], "ghost": [p[2]["n_ghost"] for p in pieces],
        }
        ok = (mism == 0 and out["force_rel"] < 1e-10 and out["force0_max_abs"] < 1e-12 and out["pe0_rel"] < 1e-9 and out["pe_trace_rel"] < 1e-9
              and out["ke_trace_rel"] < 1e-9)
        out["ok"] = bool(ok)
        print(json.dumps(out), flush=True)
    dist.barrier()
    mgr.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
