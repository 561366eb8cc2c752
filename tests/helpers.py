"""Shared helpers for the parity tests: build matching GPU managers and oracle objects."""
from __future__ import annotations

import numpy as np

from oracle.pis_oracle import Oracle
from pis_b200 import Atoms, LennardJones, LJCudaManager, SimulationBox
from pis_b200.lattice import ARGON, fcc_argon

SIGMA = ARGON["sigma"]
RC25 = 2.5 * SIGMA          # 8.5125
SKIN = 0.3 * SIGMA          # 1.0215


def argon_pair(rc=RC25):
    return LennardJones(ARGON["epsilon"], ARGON["sigma"], rc, True)


# kernel-variant combinations (force_variant, build_variant, cell_div); see DESIGN.md
VARIANTS = {
    0: None,              # library defaults
    1: (1, 1, 1),         # v1: plain all-FP64 kernels, reference-sized cells
    2: (2, 2, 1),         # v2: FP32 pre-filter + queue force kernel, pre-filter build
    3: (3, 2, 1),         # v3: 4-wide prefetching all-FP64 force kernel, pre-filter build
    4: (3, 2, 2),         # v3 + half-size cells (5^3 stencil)
    6: (3, 3, 1),         # v3 force + v3 build (packed-FP32 pair records, bit-mask append) == the defaults, explicit
    7: (3, 3, 2),         # v3 build with half-size cells (5^3 stencil)
    8: (6, 3, 2),         # k_force_split<8>: 8 lanes per atom taking whole K-tiles (round-1 small-system kernel)
    11: (7, 3, 2),        # k_force_q: four lanes per atom, lane l takes entry l of every K-tile (the automatic choice up to 75k atoms)
    12: (7, 3, 1),        # k_force_q over reference-sized cells
}


def make_manager(skin=0.0, rc=RC25, table=None, variant=0):
    m = LJCudaManager(skin=skin)
    if VARIANTS[variant]:
        fv, bv, cd = VARIANTS[variant]
        m.set_option("force_variant", fv)
        m.set_option("build_variant", bv)
        m.set_option("cell_div", cd)
    if table is None:
        m.insert((1, 1), argon_pair(rc))
    else:
        for k, p in table.items():
            m.insert(k, p)
    return m


def make_oracle(atoms: Atoms, table: dict):
    b = atoms.sim_box
    nt = max(len(atoms.masses), max(max(k) for k in table))
    masses = np.zeros(nt)
    masses[: len(atoms.masses)] = atoms.masses
    o = Oracle(b.h_colmajor(), pbc=[int(p) for p in b.pbc], n_types=nt, masses=masses,
               shift=all(p.shift for p in table.values()))
    for (i, j), p in table.items():
        o.insert(i, j, p.epsilon, p.sigma, p.rcut)
    return o


def force_rel_err(f_gpu, f_ref):
    """|dF_i| / max(|F_i|, 1e-3 * rms|F|)  (SURVEY 'hard part 4': lattice forces cancel to ~1e-14)."""
    mag = np.linalg.norm(f_ref, axis=1)
    rms = np.sqrt((mag ** 2).mean())
    den = np.maximum(mag, 1e-3 * rms if rms > 0 else 1.0)
    return np.linalg.norm(f_gpu - f_ref, axis=1) / den


def csr_rows_sorted(start, nbr):
    return [np.sort(nbr[start[i]:start[i + 1]]) for i in range(len(start) - 1)]
