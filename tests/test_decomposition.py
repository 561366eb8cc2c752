"""Host-side logic of the multi-GPU path on CPU: brick geometry, ownership, the ghost rule, and the
decomposition-independent input generator -- including a world_size-2 gloo run of the plumbing."""
import os
import socket

import numpy as np
import pytest

from pis_b200.decomposition import (brick_bounds, brick_coords, create_velocities_distributed, fcc_brick, gaussian_by_id,
                                    ghost_destinations, grid_for, owner_rank)
from pis_b200.lattice import ARGON, fcc_positions


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_bricks_partition_the_lattice(world):
    grid = grid_for(world)
    nc = 4
    L = [nc * ARGON["a"]] * 3
    allpos = fcc_positions(ARGON["a"], nc, nc, nc)
    seen = np.zeros(len(allpos), dtype=int)
    for r in range(world):
        pos, gid = fcc_brick(nc, r, grid)
        assert np.array_equal(allpos[gid], pos)
        assert (owner_rank(pos, L, grid) == r).all()
        lo, hi = brick_bounds(r, grid, L)
        assert ((pos >= lo - 1e-12) & (pos < hi + 1e-12)).all()
        seen[gid] += 1
        assert brick_coords(r, grid) == (r % grid[0], (r // grid[0]) % grid[1], r // (grid[0] * grid[1]))
    assert (seen == 1).all()
    with pytest.raises(ValueError):
        grid_for(3)


def test_owner_rank_wraps_unwrapped_positions():
    L = [20.0, 20.0, 20.0]
    pos = np.array([[1.0, 1.0, 1.0], [11.0, 1.0, 1.0], [-1.0, 21.0, 19.0], [41.0, -19.0, 10.0]])
    assert owner_rank(pos, L, (2, 2, 2)).tolist() == [0, 1, 0b101, 0b100]


def test_ghost_rule_covers_every_cross_brick_pair():
    """Every pair within the list cutoff whose atoms live on different ranks must find the partner among the
    ghosts the rule sends (checked by brute force with the minimum image)."""
    rng = np.random.default_rng(3)
    L = np.array([40.0, 36.0, 44.0])
    gw = 7.5
    pos = rng.uniform(0, 1, size=(700, 3)) * L
    for world in (2, 4, 8):
        grid = grid_for(world)
        own = owner_rank(pos, L, grid)
        have = {r: set(np.nonzero(own == r)[0].tolist()) for r in range(world)}
        for r in range(world):
            mine = np.nonzero(own == r)[0]
            for dst, idx in ghost_destinations(pos[mine], r, grid, L, gw).items():
                assert dst != r
                have[dst] |= set(mine[idx].tolist())
        d = pos[None] - pos[:, None]
        d -= L * np.round(d / L)
        close = np.sqrt((d * d).sum(-1)) <= gw * 0.999
        for i, j in zip(*np.nonzero(close)):
            assert j in have[own[i]], (world, i, j)


def test_gaussian_by_id_is_decomposition_independent():
    full = gaussian_by_id(np.arange(5000), 42)
    part = gaussian_by_id(np.arange(1234, 2345), 42)
    assert np.array_equal(full[1234:2345], part)
    assert abs(full.mean()) < 0.03 and abs(full.std() - 1.0) < 0.03
    assert not np.array_equal(full, gaussian_by_id(np.arange(5000), 43))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from pis_b200.distributed import allreduce_sum_host, gather_by_gid, init_process_group

    r, w = init_process_group()
    assert (r, w) == (rank, world)
    grid = grid_for(world)
    nc = 6
    n_global = 4 * nc ** 3
    pos, gid = fcc_brick(nc, rank, grid)
    m = np.full(len(gid), ARGON["mass"])
    vel = create_velocities_distributed(gid, m, 25.0, 9, n_global, allreduce_sum_host)
    (V, X) = gather_by_gid(gid, [vel, pos], n_global)
    if rank == 0:
        np.save(os.path.join(out_dir, "V.npy"), V)
        np.save(os.path.join(out_dir, "X.npy"), X)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_plumbing(tmp_path):
    """Two CPU processes over gloo: brick generation + globally reduced velocity post-processing + gather
    reproduce the single-process arrays (the N>1 host path, without GPUs)."""
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    V = np.load(tmp_path / "V.npy")
    X = np.load(tmp_path / "X.npy")
    nc = 6
    n_global = 4 * nc ** 3
    assert np.array_equal(X, fcc_positions(ARGON["a"], nc, nc, nc))
    ref = create_velocities_distributed(np.arange(n_global), np.full(n_global, ARGON["mass"]), 25.0, 9, n_global)
    assert np.abs(V - ref).max() < 1e-13
    ke = 0.5 * ARGON["mass"] * (V * V).sum()
    assert abs(2 * ke / (3 * n_global * 0.0083144621) - 25.0) < 1e-9
