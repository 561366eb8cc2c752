"""Synthetic inputs: FCC supercells in the reference generator's atom order
(src/bin/sapphire/main.rs:17-24,61-77: ix -> iy -> iz -> basis) and `velocity ... create`
post-processing (drift removal + rescale, src/atoms/velocities.rs:35-59).

The reference's Gaussian stream (rand 0.10 SmallRng + rand_distr Normal) is third-party and
unpinned, so velocities come from a generator defined here (splitmix64 + Box-Muller keyed by the
global atom id, see decomposition.gaussian_by_id); the same
arrays are then fed to both the GPU path and the oracle.
"""
from __future__ import annotations

import numpy as np

from .atoms import KB_KJPERMOLEKELVIN, Atoms
from .simulation_box import SimulationBox

ARGON = dict(epsilon=0.238, sigma=3.405, mass=39.948, a=5.41)  # example/argon4000.txt:6-14

FCC_BASIS = np.array([[0.0, 0.0, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]])


def fcc_positions(a: float, nx: int, ny: int, nz: int) -> np.ndarray:
    """pos = h * (cell_origin + basis) with h = a*I, atom order ix, iy, iz, basis."""
    ix, iy, iz = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64),
                             np.arange(nz, dtype=np.float64), indexing="ij")
    origin = np.stack([ix, iy, iz], axis=-1).reshape(-1, 1, 3)
    frac = origin + FCC_BASIS.reshape(1, 4, 3)
    return np.ascontiguousarray((a * frac).reshape(-1, 3))


def create_velocities(n: int, masses_per_atom: np.ndarray, temperature: float, seed: int) -> np.ndarray:
    """Gaussian velocities with sigma_i = sqrt(kB T / m_i) (velocities.rs:24), then remove_drift
    (:35-50) and rescale_to_temperature (:52-59)."""
    from .decomposition import gaussian_by_id  # counter-based: identical values on any GPU count

    sig = np.sqrt(KB_KJPERMOLEKELVIN * temperature / masses_per_atom)
    v = gaussian_by_id(np.arange(n), seed) * sig[:, None]
    total_mass = masses_per_atom.sum()
    vcm = (v * masses_per_atom[:, None]).sum(axis=0) / total_mass
    v -= vcm[None, :]
    ke = float((0.5 * masses_per_atom * (v * v).sum(axis=1)).sum())
    t_cur = (2.0 * ke) / (3.0 * n * KB_KJPERMOLEKELVIN)
    if t_cur > 0.0:
        v *= np.sqrt(temperature / t_cur)
    return np.ascontiguousarray(v)


def fcc_argon(ncell: int, temperature: float = 5.0, seed: int = 12345, jitter: float = 0.0, pinned: bool = False,
              a: float | None = None) -> Atoms:
    """ncell^3 x 4 argon atoms on an FCC lattice (a = 5.41 A), optional Gaussian position jitter (A)."""
    a = ARGON["a"] if a is None else a
    pos = fcc_positions(a, ncell, ncell, ncell)
    n = pos.shape[0]
    L = ncell * a
    if jitter > 0.0:
        rng = np.random.Generator(np.random.Philox(seed + 1))
        pos = pos + rng.standard_normal(pos.shape) * jitter
        pos -= np.floor(pos / L) * L
        pos = np.where(pos >= L, pos - L, pos)
    box = SimulationBox.from_lammps_data(0.0, L, 0.0, L, 0.0, L)
    types = np.ones(n, dtype=np.int32)
    m = np.full(n, ARGON["mass"])
    vel = create_velocities(n, m, temperature, seed) if temperature > 0.0 else np.zeros((n, 3))
    return Atoms(types, [ARGON["mass"]], pos, box, velocities=vel, pinned=pinned)
