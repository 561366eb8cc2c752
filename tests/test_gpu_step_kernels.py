"""k_force_q -- four lanes per atom, lane l takes entry l of every K-tile, integrator in the epilogue (the automatic step
kernel up to 75k atoms; force_variant 7 forces it) -- and, beside it, the fused thread-per-atom kernel forced at small sizes
(force_variant 3).  Same per-atom lists, same lean arithmetic, so the oracle bars apply unchanged; exercised on the shapes
where lane groups are ragged (odd counts, vacuum, grids of a few cells, very short and very long rows, two types)."""
import numpy as np
import pytest

from pis_b200 import Atoms, SimulationBox
from pis_b200.lattice import fcc_argon
from tests.helpers import SKIN, argon_pair, csr_rows_sorted, force_rel_err, make_manager, make_oracle

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-10
ENERGY_TOL = 1e-9


def _check_against_oracle(atoms, variant=9, rc=None, steps=30):
    table = {(1, 1): argon_pair(rc) if rc else argon_pair()}
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN)
    mgr = make_manager(skin=SKIN, variant=variant, table=table)
    mgr.attach(atoms)
    rows = mgr.neighbours(atoms.n_atoms)
    for i, (a_, b_) in enumerate(zip(rows, csr_rows_sorted(start, nbr))):
        assert np.array_equal(a_, b_), f"atom {i}"
    pe = mgr.compute()
    mgr.download(atoms, positions=False, velocities=False)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
    if steps:
        x, v, f = atoms.positions.copy(), atoms.velocities.copy(), f_ref.copy()
        pes, kes = [], []
        for _ in range(steps):
            pes.append(orc.verlet_step_nve(x, v, f, atoms.type_ids, 0.25))
            kes.append(orc.kinetic_energy(v, atoms.type_ids))
        th = mgr.step_nve(0.25, steps)
        assert np.max(np.abs(th["pe"] - np.array(pes)) / np.abs(pes)) <= ENERGY_TOL
        assert np.max(np.abs(th["ke"] - np.array(kes)) / np.abs(kes)) <= ENERGY_TOL
        mgr.download(atoms)
        assert np.abs(atoms.positions - x).max() <= 1e-9
        assert force_rel_err(atoms.forces, f).max() <= FORCE_TOL
    return mgr



@pytest.mark.parametrize("variant", [6, 11, 12])
@pytest.mark.parametrize("ncell", [6, 7, 11])
def test_step_kernel_variants_match_the_oracle(ncell, variant):
    """ncell = 6: 32.5 A box, the smallest the reference handles without double counting (6 half-size cells per edge, every
    warp needs the minimum image); 11: 59.5 A, interior warps exist."""
    _check_against_oracle(fcc_argon(ncell, temperature=40.0, seed=ncell, jitter=0.2), variant=variant)


@pytest.mark.parametrize("variant", [6, 11])
def test_step_kernel_variants_odd_atom_count_vacuum_slab_and_non_cubic_box(variant):
    """Half the box empty (empty cells, bricks with no atoms, uneven density), three different box edges, an odd atom count."""
    base = fcc_argon(10, temperature=25.0, seed=8, jitter=0.1)
    a = 5.41
    keep = np.flatnonzero(base.positions[:, 0] < 5 * a)
    if len(keep) % 2 == 0:
        keep = keep[:-1]
    box = SimulationBox.from_lammps_data(0, 10 * a + 2.1, 0, 10 * a + 3.7, 0, 10 * a + 9.1)
    atoms = Atoms(np.ones(len(keep), dtype=np.int32), [39.948], base.positions[keep].copy(), box, velocities=base.velocities[keep].copy())
    _check_against_oracle(atoms, variant=variant)


@pytest.mark.parametrize("variant", [6, 11])
def test_step_kernel_variants_long_cutoff(variant):
    """rc = 4 sigma: ~330 entries per row (83 K-tiles), capacity regrown."""
    atoms = fcc_argon(10, temperature=40.0, seed=3, jitter=0.12)
    mgr = _check_against_oracle(atoms, variant=variant, rc=4.0 * 3.405, steps=10)
    assert mgr.stats()["max_neighbours"] > 250


@pytest.mark.parametrize("variant", [6, 11])
def test_step_kernel_variants_hot_run_with_rebuilds(variant):
    """300 steps at 60 K (a rebuild every few steps): PE / KE traces vs the oracle; 13 x 13 x 13 half-size cells."""
    atoms = fcc_argon(12, temperature=60.0, seed=21)
    orc = make_oracle(atoms, {(1, 1): argon_pair(8.5)})
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref = orc.run_nve(x, v, np.zeros_like(x), atoms.type_ids, 0.25, 300)
    mgr = make_manager(skin=SKIN, rc=8.5, variant=variant)
    mgr.attach(atoms)
    pe0 = mgr.compute()
    assert abs(pe0 - ref[0, 0]) <= ENERGY_TOL * abs(ref[0, 0])
    th = mgr.step_nve(0.25, 300)
    assert mgr.stats()["n_builds"] >= 10
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL
    mgr.download(atoms)
    assert np.abs(atoms.positions - x).max() < 1e-8


def test_four_lanes_two_types_and_a_missing_pair():
    """The type-table form of the step kernels (MULTI): per-pair cutoffs, a missing (2,2) entry skipped like lennard_jones.rs:216-222."""
    from pis_b200 import LennardJones

    table = {(1, 1): LennardJones(0.238, 3.405, 8.5), (1, 2): LennardJones(0.15, 3.0, 7.5)}
    for variant in (11, 6):
        atoms = fcc_argon(8, temperature=30.0, seed=9, jitter=0.1)
        atoms.type_ids[::3] = 2
        atoms.masses = [39.948, 20.18]
        orc = make_oracle(atoms, table)
        orc.lib.orc_set_quiet(1)
        pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
        mgr = make_manager(skin=SKIN, variant=variant, table=table)
        pe = mgr.compute_potential(atoms)
        assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
        assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
        x, v, f = atoms.positions.copy(), atoms.velocities.copy(), f_ref.copy()
        pes = [orc.verlet_step_nve(x, v, f, atoms.type_ids, 0.25) for _ in range(30)]
        mgr.attach(atoms)
        mgr.compute()
        th = mgr.step_nve(0.25, 30)
        assert np.max(np.abs(th["pe"] - np.array(pes)) / np.abs(pes)) <= ENERGY_TOL
        orc.lib.orc_set_quiet(0)
