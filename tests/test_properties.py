"""Property tests (plan iii), hypothesis-driven.

CPU half (`-m "not gpu"`): the oracle -- the restatement of lennard_jones.rs:186-244 / :345-415 -- on random small systems in
random orthorhombic boxes: Newton's third law, translation and permutation invariance, the cell driver against a brute-force
numpy evaluation, list symmetry.  GPU half (`-m gpu`): the CUDA path through the C ABI against the oracle on the same random
draws (neighbour sets exact, forces 1e-10), list-with-skin == no-list forces, and rebuild-trigger safety: after any number
of steps every pair inside the cutoff is in the list the step kernel used."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from pis_b200 import Atoms, LennardJones, SimulationBox
from tests.helpers import csr_rows_sorted, force_rel_err, make_manager, make_oracle

SIGMA, EPS = 3.405, 0.238


def _system(seed, n, lx, ly, lz, rc, two_types=False):
    """n atoms at random positions, no two closer than 0.8 sigma (rejection), in an lx x ly x lz box (edges >= 3 rc)."""
    rng = np.random.default_rng(seed)
    L = np.array([lx, ly, lz])
    pos = np.empty((0, 3))
    while len(pos) < n:
        cand = rng.random((4 * n, 3)) * L
        for c in cand:
            if len(pos) == 0:
                pos = c[None, :]
                continue
            d = pos - c
            d -= np.round(d / L) * L
            if (np.einsum("ij,ij->i", d, d) > (0.8 * SIGMA) ** 2).all():
                pos = np.vstack([pos, c])
                if len(pos) == n:
                    break
    types = np.ones(n, dtype=np.int32)
    masses = [39.948]
    table = {(1, 1): LennardJones(EPS, SIGMA, rc, True)}
    if two_types:
        types[rng.random(n) < 0.4] = 2
        masses = [39.948, 83.798]
        table = {(1, 1): LennardJones(EPS, SIGMA, rc, True), (1, 2): LennardJones(0.3, 3.5, 0.9 * rc, True),
                 (2, 2): LennardJones(0.4, 3.6, 0.8 * rc, True)}
    box = SimulationBox.from_lammps_data(0.0, lx, 0.0, ly, 0.0, lz)
    vel = rng.standard_normal((n, 3)) * 0.02
    return Atoms(types, masses, np.ascontiguousarray(pos), box, velocities=vel), table, L


def _brute_force(pos, L, types, table):
    """O(N^2) numpy evaluation with a different minimum-image formula (d - L round(d / L)), pair terms of lennard_jones.rs:33-55."""
    n = len(pos)
    f = np.zeros((n, 3))
    pe = 0.0
    for i in range(n):
        d = pos - pos[i]
        d -= np.round(d / L) * L
        r2 = np.einsum("ij,ij->i", d, d)
        for j in range(i + 1, n):
            p = table.get((min(types[i], types[j]), max(types[i], types[j])))
            if p is None or np.sqrt(r2[j]) > p.rcut:
                continue
            inv = 1.0 / r2[j]
            s6 = (p.sigma ** 2 * inv) ** 3
            s12 = s6 * s6
            ucut = 4.0 * p.epsilon * ((p.sigma / p.rcut) ** 12 - (p.sigma / p.rcut) ** 6)
            pe += 4.0 * p.epsilon * (s12 - s6) - ucut
            fv = 24.0 * p.epsilon * (2.0 * s12 - s6) * inv * d[j]
            f[i] -= fv
            f[j] += fv
    return pe, f


box_edge = st.floats(min_value=27.0, max_value=40.0)
# derandomize: the same examples in every run (the driver's suite must not depend on a seed); widen locally with
# HYPOTHESIS_PROFILE-style edits when hunting
common = dict(deadline=None, derandomize=True, database=None,
              suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


@settings(max_examples=25, **common)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(8, 90), lx=box_edge, ly=box_edge, lz=box_edge, two=st.booleans())
def test_oracle_cell_driver_equals_brute_force_and_obeys_the_symmetries(seed, n, lx, ly, lz, two):
    atoms, table, L = _system(seed, n, lx, ly, lz, rc=8.5, two_types=two)
    orc = make_oracle(atoms, table)
    pe, f = orc.compute_potential(atoms.positions, atoms.type_ids)
    pe_b, f_b = _brute_force(atoms.positions, L, atoms.type_ids, table)
    scale = max(np.abs(f_b).max(), 1e-12)
    assert abs(pe - pe_b) <= 1e-11 * max(abs(pe_b), 1.0)
    assert np.abs(f - f_b).max() <= 1e-11 * scale
    assert np.abs(f.sum(axis=0)).max() <= 1e-10 * scale                      # Newton's third law
    shift = np.array([0.37 * lx, -1.2 * ly, 2.0 * lz + 0.11])                 # translation (re-wrapped)
    pos2 = atoms.positions + shift
    pos2 -= np.floor(pos2 / L) * L
    pe2, f2 = orc.compute_potential(np.ascontiguousarray(pos2), atoms.type_ids)
    assert abs(pe2 - pe) <= 1e-10 * max(abs(pe), 1.0)
    assert np.abs(f2 - f).max() <= 1e-9 * scale
    perm = np.random.default_rng(seed + 1).permutation(n)                    # permutation of the atom order
    pe3, f3 = orc.compute_potential(np.ascontiguousarray(atoms.positions[perm]), np.ascontiguousarray(atoms.type_ids[perm]))
    assert abs(pe3 - pe) <= 1e-11 * max(abs(pe), 1.0)
    assert np.abs(f3 - f[perm]).max() <= 1e-11 * scale


@settings(max_examples=20, **common)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(8, 90), lx=box_edge, ly=box_edge, lz=box_edge, skin=st.floats(0.0, 0.5))
def test_oracle_neighbour_list_is_symmetric_and_complete(seed, n, lx, ly, lz, skin):
    atoms, table, L = _system(seed, n, lx, ly, lz, rc=8.0)
    orc = make_oracle(atoms, table)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=skin)
    rows = [set(nbr[start[i]:start[i + 1]].tolist()) for i in range(n)]
    for i in range(n):
        d = atoms.positions - atoms.positions[i]
        d -= np.round(d / L) * L
        r = np.sqrt(np.einsum("ij,ij->i", d, d))
        expect = {j for j in range(n) if j != i and r[j] <= 8.0 + skin}
        near_edge = {j for j in range(n) if j != i and abs(r[j] - (8.0 + skin)) < 1e-9}
        assert rows[i] - near_edge == expect - near_edge
        for j in rows[i]:
            assert i in rows[j]


# ---------------------------------------------------------------------------------------------------------------------
# GPU half
# ---------------------------------------------------------------------------------------------------------------------
gpu_box_edge = st.floats(min_value=40.0, max_value=64.0)


@pytest.mark.gpu
@settings(max_examples=20, **common)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(50, 1500), lx=gpu_box_edge, ly=gpu_box_edge, lz=gpu_box_edge,
       two=st.booleans(), skin=st.sampled_from([0.0, 0.4, 1.0215]), variant=st.sampled_from([0, 1, 4, 7, 11]))
def test_gpu_matches_the_oracle_on_random_systems(seed, n, lx, ly, lz, two, skin, variant):
    """Random gas-like configurations (empty cells, ragged rows, interior and boundary warps, per-pair cutoffs) through
    every kernel family: neighbour sets exact, forces 1e-10, PE 1e-9, and list-with-skin == no-list forces."""
    atoms, table, L = _system(seed, n, lx, ly, lz, rc=8.5, two_types=two)
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=skin)
    mgr = make_manager(skin=skin, table=table, variant=variant)
    try:
        mgr.attach(atoms)
        rows = mgr.neighbours(n)
        for i, (a, b) in enumerate(zip(rows, csr_rows_sorted(start, nbr))):
            assert np.array_equal(a, b), f"atom {i}"
        pe = mgr.compute()
        mgr.download(atoms, positions=False, velocities=False)
        assert abs(pe - pe_ref) <= 1e-9 * max(abs(pe_ref), 1e-6)
        assert force_rel_err(atoms.forces, f_ref).max() <= 1e-10
    finally:
        mgr.close()


@pytest.mark.gpu
@settings(max_examples=8, **common)
@given(seed=st.integers(0, 10 ** 6), steps=st.integers(1, 60), temp=st.floats(20.0, 200.0))
def test_rebuild_trigger_safety(seed, steps, temp):
    """After ANY number of steps at any temperature the list the next force pass would use still holds every pair inside
    the cutoff (the skin trigger fired in time): compare the device's list at that moment with a fresh brute-force search."""
    from pis_b200.lattice import fcc_argon
    from scipy.spatial import cKDTree

    atoms = fcc_argon(8, temperature=temp, seed=seed % 1000, jitter=0.1)
    rc, skin = 8.5125, 1.0215
    mgr = make_manager(skin=skin)
    try:
        mgr.attach(atoms)
        mgr.compute()
        mgr.step_nve(0.25, steps)
        builds = mgr.stats()["n_builds"]
        rows = mgr.neighbours(atoms.n_atoms)           # the list as it stands (pisb_neighbours does not rebuild a valid list)
        assert mgr.stats()["n_builds"] == builds
        mgr.download(atoms)
        L = float(atoms.sim_box.h[0, 0])
        pos = atoms.positions - np.floor(atoms.positions / L) * L
        pos = np.where(pos >= L, pos - L, pos)
        pairs = cKDTree(pos, boxsize=L).query_pairs(rc, output_type="ndarray")
        for i, j in pairs:
            assert j in rows[i] and i in rows[j], f"pair ({i}, {j}) is inside the cutoff but not listed after {steps} steps"
    finally:
        mgr.close()
