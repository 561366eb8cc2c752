"""One-process-per-GPU driver of the spatially decomposed path.  torch.distributed is the plumbing
(rendezvous, broadcast of the NCCL unique id, host-side sums); the halo exchange, migration and
thermo reductions run inside libpisb200 on its own NCCL communicator (csrc/pisb_multi.cuh)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .atoms import Atoms
from .decomposition import grid_for
from .potentials import LennardJones, LJCudaManager


def init_process_group():
    """Initialise torch.distributed from the torchrun environment (idempotent). Returns (rank, world)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        backend = "cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=int(os.environ.get("RANK", "0")),
                                world_size=int(os.environ.get("WORLD_SIZE", "1")))
    return dist.get_rank(), dist.get_world_size()


def allreduce_sum_host(x: np.ndarray) -> np.ndarray:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


class DistributedLJ(LJCudaManager):
    """LJCudaManager whose atoms are spread over the ranks of a torch.distributed job."""

    def __init__(self, skin: float, local_device: int, rank: int, world: int, grid=None):
        super().__init__(skin=skin, device=local_device)
        self.rank, self.world = rank, world
        self.grid = tuple(grid) if grid is not None else grid_for(world)
        self._comm_ready = False

    def _init_comm(self):
        import torch
        import torch.distributed as dist

        lib = capi.load()
        buf = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            capi.check(None, lib.pisb_comm_unique_id(capi._ptr(buf), 128))
        t = torch.from_numpy(buf)
        dist.broadcast(t, src=0)
        grid = np.array(self.grid, dtype=np.int32)
        capi.check(self._h, lib.pisb_comm_init(self._h, self.rank, self.world, capi._ptr(buf), capi._ptr(grid)))
        self._comm_ready = True

    def attach_owned(self, atoms: Atoms, global_ids: np.ndarray):
        """Upload this rank's atoms (global box in atoms.sim_box) with their global ids."""
        self._ensure_handle(atoms)
        if not self._comm_ready:
            self._init_comm()
        self._ensure_box(atoms)
        gids = np.ascontiguousarray(global_ids, dtype=np.int32)
        rc = capi.load().pisb_upload_owned(self._h, atoms.n_atoms, capi._ptr(atoms.positions), capi._ptr(atoms.velocities),
                                           capi._ptr(atoms.forces), capi._ptr(atoms.type_ids), capi._ptr(gids))
        capi.check(self._h, rc)
        self._atoms_id = None

    def download_owned(self, positions=True, velocities=True, forces=True):
        """(gids, pos, vel, force) of the atoms this rank owns now (row order arbitrary; gids identify the rows).
        The destination arrays are pinned and reused between calls."""
        cap = int(self.stats()["n_atoms"]) + 16
        if getattr(self, "_dl_cap", 0) < cap:
            self._dl_cap = int(cap * 1.1) + 1024
            self._dl = [capi.pinned_empty((self._dl_cap, 3)) for _ in range(3)] + [capi.pinned_empty((self._dl_cap,), np.int32)]
        pos, vel, frc, gid = self._dl
        n = C.c_int64()
        capi.check(self._h, capi.load().pisb_download_owned(self._h, self._dl_cap, capi._ptr(pos) if positions else None,
                                                            capi._ptr(vel) if velocities else None,
                                                            capi._ptr(frc) if forces else None, capi._ptr(gid), C.byref(n)))
        k = n.value
        return gid[:k], pos[:k], vel[:k], frc[:k]

    def download_owned_begin(self, positions=True, velocities=True, forces=True):
        """Asynchronous form: returns the same views at once; they hold the frame after download_end().  A second set of
        pinned destination arrays, so a synchronous download_owned between begin and end does not disturb the frame."""
        cap = int(self.stats()["n_atoms"]) + 16
        if getattr(self, "_dla_cap", 0) < cap:
            self._dla_cap = int(cap * 1.1) + 1024
            self._dla = [capi.pinned_empty((self._dla_cap, 3)) for _ in range(3)] + [capi.pinned_empty((self._dla_cap,), np.int32)]
        pos, vel, frc, gid = self._dla
        n = C.c_int64()
        capi.check(self._h, capi.load().pisb_download_owned_begin(self._h, self._dla_cap, capi._ptr(pos) if positions else None,
                                                                  capi._ptr(vel) if velocities else None,
                                                                  capi._ptr(frc) if forces else None, capi._ptr(gid), C.byref(n)))
        k = n.value
        return gid[:k], pos[:k], vel[:k], frc[:k]

    def neighbours_owned(self):
        """(gids, rows): sorted global-id neighbour rows of the owned atoms and the global id of each row (test hook)."""
        n = int(self.stats()["n_atoms"])
        rows = self.neighbours(n)
        gid = np.zeros(n + 16, dtype=np.int32)
        k = C.c_int64()
        capi.check(self._h, capi.load().pisb_owned_ids(self._h, n + 16, capi._ptr(gid), C.byref(k)))
        assert k.value == n
        return gid[:n].copy(), rows


def gather_by_gid(gid, arrays, n_global):
    """Assemble per-rank (gid, array...) pieces into global arrays on every rank (tests / dumps)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    pieces = [None] * world
    dist.all_gather_object(pieces, (gid, arrays))
    outs = [np.zeros((n_global,) + a.shape[1:], dtype=a.dtype) for a in arrays]
    seen = np.zeros(n_global, dtype=np.int32)
    for g, arrs in pieces:
        seen[g] += 1
        for o, a in zip(outs, arrs):
            o[g] = a
    assert (seen == 1).all(), "every atom must be owned by exactly one rank"
    return outs
