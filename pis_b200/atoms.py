"""Atoms -- host mirror of `struct Atoms` (src/atoms/new.rs:9-17).

positions / velocities / forces are (N, 3) C-contiguous float64 arrays: the same memory image as
nalgebra's column-major Matrix3xX<f64> (xyz-interleaved), so the same pointer can be handed to the
C ABI.  type_ids are 1-based; masses are indexed by type-1 (src/atoms/properties.rs:9-13).
The per-step observables are computed on the device (pisb_thermo); the scalar formulas of
src/atoms/properties.rs that turn them into T and P live here.
"""
from __future__ import annotations

import numpy as np

from .simulation_box import SimulationBox

KB_KJPERMOLEKELVIN = 0.0083144621  # src/constants.rs:3


class Atoms:
    def __init__(self, type_ids, masses, positions, sim_box: SimulationBox, velocities=None, forces=None,
                 pinned: bool = False):
        n = len(type_ids)
        self.n_atoms = n
        self.type_ids = np.ascontiguousarray(type_ids, dtype=np.int32)
        self.masses = [float(m) for m in masses]
        alloc = self._alloc_pinned if pinned else (lambda: np.zeros((n, 3), dtype=np.float64))
        self.positions = alloc()
        self.positions[...] = np.asarray(positions, dtype=np.float64).reshape(n, 3)
        self.velocities = alloc()
        if velocities is not None:
            self.velocities[...] = np.asarray(velocities, dtype=np.float64).reshape(n, 3)
        self.forces = alloc()  # reader zero-initialises (commands.rs:361)
        if forces is not None:
            self.forces[...] = np.asarray(forces, dtype=np.float64).reshape(n, 3)
        self.sim_box = sim_box

    def _alloc_pinned(self):
        from .capi import pinned_empty

        a = pinned_empty((self.n_atoms, 3))
        a[...] = 0.0
        return a

    def mass_i(self, i: int) -> float:
        """ref: src/atoms/properties.rs:9-13"""
        return self.masses[int(self.type_ids[i]) - 1]

    def degress_of_freedom(self) -> int:
        """ref: src/atoms/properties.rs:41-43 (3N, no drift correction)"""
        return 3 * self.n_atoms

    def temerature(self, kinetic_energy: float) -> float:
        """ref: src/atoms/properties.rs:28-30 (name kept as in the reference)"""
        return (2.0 * kinetic_energy) / (float(self.degress_of_freedom()) * KB_KJPERMOLEKELVIN)

    temperature = temerature

    def pressure(self, kinetic_energy: float, virial_trace: float) -> float:
        """ref: src/atoms/properties.rs:61-65 with virial = tr(X F^T) from the device"""
        return (2.0 * kinetic_energy + virial_trace) / (3.0 * self.sim_box.volume())
