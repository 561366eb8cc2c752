"""Host-side logic of the multi-GPU spatial decomposition (pure numpy: testable on CPU under gloo).

One process per GPU.  The periodic box is cut into P[0] x P[1] x P[2] bricks with P_d in {1, 2}
(1, 2x1x1, 2x2x1, 2x2x2: one NVSwitch node, all peers equidistant, so the shape only minimises
surface).  rank = (bz * Py + by) * Px + bx.  The device code (csrc/pisb_multi.cuh) uses the same
formulas; tests/test_decomposition.py checks them against each other through these functions.
"""
from __future__ import annotations

import numpy as np

from .atoms import KB_KJPERMOLEKELVIN
from .lattice import ARGON, FCC_BASIS


def grid_for(nranks: int) -> tuple[int, int, int]:
    try:
        return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[nranks]
    except KeyError:
        raise ValueError("supported GPU counts on one node: 1, 2, 4, 8") from None


def brick_coords(rank: int, grid) -> tuple[int, int, int]:
    return rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1])


def brick_bounds(rank: int, grid, box_lengths):
    b = brick_coords(rank, grid)
    lo = np.array([b[d] * (box_lengths[d] / grid[d]) for d in range(3)])
    hi = np.array([box_lengths[d] if b[d] == grid[d] - 1 else (b[d] + 1) * (box_lengths[d] / grid[d]) for d in range(3)])
    return lo, hi


def owner_rank(pos: np.ndarray, box_lengths, grid) -> np.ndarray:
    """Rank owning each (possibly unwrapped) position: floor(frac(x / L) * P), like dest_rank on the device."""
    r = np.zeros(pos.shape[0], dtype=np.int64)
    mul = 1
    for d in range(3):
        s = pos[:, d] * (1.0 / box_lengths[d])
        s = s - np.floor(s)
        bd = np.clip(np.floor(s * grid[d]).astype(np.int64), 0, grid[d] - 1)
        r += bd * mul
        mul *= grid[d]
    return r


def ghost_destinations(pos: np.ndarray, rank: int, grid, box_lengths, gw: float):
    """Reference statement of the ghost rule: for every non-empty subset S of the decomposed dimensions, an
    owned atom within gw of a brick face in every d in S is sent to the rank with those brick coordinates
    flipped.  Returns {dest_rank: sorted atom indices}."""
    lo, hi = brick_bounds(rank, grid, box_lengths)
    b = brick_coords(rank, grid)
    near = np.zeros((pos.shape[0], 3), dtype=bool)
    for d in range(3):
        if grid[d] > 1:
            near[:, d] = (pos[:, d] - lo[d] < gw) | (hi[d] - pos[:, d] < gw)
    out: dict[int, np.ndarray] = {}
    for sset in range(1, 8):
        dims = [d for d in range(3) if sset & (1 << d)]
        if any(grid[d] == 1 for d in dims):
            continue
        sel = np.all(near[:, dims], axis=1)
        bb = list(b)
        for d in dims:
            bb[d] ^= 1
        dst = (bb[2] * grid[1] + bb[1]) * grid[0] + bb[0]
        idx = np.nonzero(sel)[0]
        if idx.size:
            out[dst] = np.union1d(out.get(dst, np.empty(0, dtype=np.int64)), idx)
    return out


# ---- decomposition-independent synthetic input --------------------------------------------------
def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def gaussian_by_id(ids: np.ndarray, seed: int) -> np.ndarray:
    """(len(ids), 3) standard normals that depend only on (seed, global id, component): counter-based
    (splitmix64 + Box-Muller), so every rank can generate exactly its own atoms' values."""
    ids = np.asarray(ids, dtype=np.uint64)
    out = np.empty((ids.size, 3))
    with np.errstate(over="ignore"):
        base = _splitmix64(np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + np.uint64(12345))
        for c in range(3):
            k = _splitmix64(ids * np.uint64(6) + np.uint64(2 * c) + base)
            k2 = _splitmix64(ids * np.uint64(6) + np.uint64(2 * c + 1) + base)
            u1 = ((k >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740993.0)  # (0, 1)
            u2 = (k2 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)         # [0, 1)
            out[:, c] = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return out


def fcc_brick(ncell, rank: int, grid, a: float | None = None):
    """Lattice sites of the (nx, ny, nz)-cell FCC supercell (ncell: int or 3-tuple) that fall in this
    rank's brick, with their GLOBAL ids in the reference generator's order
    (src/bin/sapphire/main.rs:61-66: ix, iy, iz, basis)."""
    a = ARGON["a"] if a is None else a
    nc = (ncell, ncell, ncell) if np.isscalar(ncell) else tuple(int(c) for c in ncell)
    b = brick_coords(rank, grid)
    rng = []
    for d in range(3):
        if nc[d] % grid[d]:
            raise ValueError("cells per dimension must be divisible by the bricks per dimension")
        w = nc[d] // grid[d]
        rng.append(np.arange(b[d] * w, (b[d] + 1) * w, dtype=np.int64))
    ix, iy, iz = np.meshgrid(*rng, indexing="ij")
    cell = np.stack([ix, iy, iz], axis=-1).reshape(-1, 1, 3)
    pos = (a * (cell.astype(np.float64) + FCC_BASIS.reshape(1, 4, 3))).reshape(-1, 3)
    cid = ((cell[:, 0, 0] * nc[1] + cell[:, 0, 1]) * nc[2] + cell[:, 0, 2]).reshape(-1, 1)
    gid = (cid * 4 + np.arange(4, dtype=np.int64).reshape(1, 4)).reshape(-1)
    return np.ascontiguousarray(pos), gid.astype(np.int32)


def create_velocities_distributed(gids, masses_per_atom, temperature, seed, n_global, allreduce=None):
    """velocity all create T seed, decomposition-independent: Gaussian by global id, then the reference's
    remove_drift and rescale_to_temperature (src/atoms/velocities.rs:35-59) with GLOBAL sums
    (allreduce(np.ndarray) -> summed array; identity on one rank)."""
    if allreduce is None:
        allreduce = lambda x: x  # noqa: E731
    sig = np.sqrt(KB_KJPERMOLEKELVIN * temperature / masses_per_atom)
    v = gaussian_by_id(gids, seed) * sig[:, None]
    sums = allreduce(np.concatenate([[masses_per_atom.sum()], (v * masses_per_atom[:, None]).sum(axis=0)]))
    v -= (sums[1:4] / sums[0])[None, :]
    ke = allreduce(np.array([(0.5 * masses_per_atom * (v * v).sum(axis=1)).sum()]))[0]
    t_cur = (2.0 * ke) / (3.0 * n_global * KB_KJPERMOLEKELVIN)
    if t_cur > 0.0:
        v *= np.sqrt(temperature / t_cur)
    return np.ascontiguousarray(v)
