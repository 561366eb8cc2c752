"""The two-atoms-per-thread path (pisb_pairlist.cuh; automatic above 75k atoms, forced here with force_variant 5): its
three-section lists must reassemble into exactly the oracle's per-atom neighbour sets, and its forces / traces must meet the
same bars as every other kernel -- including the shapes where a pair thread cannot sweep both atoms jointly (row ends,
vacuum gaps, odd atom counts, grids of 5 cells)."""
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


@pytest.mark.parametrize("variant", [9, 10])
@pytest.mark.parametrize("ncell", [6, 7, 11])
def test_pair_lists_match_the_oracle(ncell, variant):
    """ncell = 6: 32.5 A box, the smallest the reference handles without double counting (3 cells of rc + skin), a grid of 6
    half-size cells per edge -- the joint sweep of atoms two cells apart would lap the periodic row and falls back to two
    sweeps; 11: joint sweeps, interior warps."""
    _check_against_oracle(fcc_argon(ncell, temperature=40.0, seed=ncell, jitter=0.2), variant=variant)


def test_pair_lists_odd_atom_count_and_shuffled_ids():
    """An odd number of atoms leaves the last pair thread with one atom; ids in random order decouple slot pairs from ids."""
    base = fcc_argon(8, temperature=30.0, seed=4, jitter=0.15)
    rng = np.random.default_rng(8)
    keep = rng.permutation(base.n_atoms)[: base.n_atoms - 1]
    atoms = Atoms(np.ones(len(keep), dtype=np.int32), [39.948], base.positions[keep].copy(), base.sim_box,
                  velocities=base.velocities[keep].copy())
    _check_against_oracle(atoms)


def test_pair_lists_vacuum_slab_and_non_cubic_box():
    """Half the box empty: slot neighbours at the slab surface can be many cells apart (the pair thread sweeps each atom on
    its own); three different box edges."""
    base = fcc_argon(8, temperature=25.0, seed=8, jitter=0.1)
    a = 5.41
    keep = base.positions[:, 0] < 4 * a                      # vacuum in x: the direction slot order runs along
    box = SimulationBox.from_lammps_data(0, 8 * a + 2.1, 0, 8 * a + 3.7, 0, 8 * a + 9.1)
    atoms = Atoms(np.ones(keep.sum(), dtype=np.int32), [39.948], base.positions[keep].copy(), box, velocities=base.velocities[keep].copy())
    _check_against_oracle(atoms)


def test_pair_lists_long_cutoff_regrows_capacity():
    """rc = 4 sigma at rho* ~ 1: ~330 neighbours per atom, beyond the first capacity estimate; sections regrow with the list."""
    atoms = fcc_argon(10, temperature=40.0, seed=3, jitter=0.12)
    mgr = _check_against_oracle(atoms, rc=4.0 * 3.405, steps=10)
    st = mgr.stats()
    assert st["max_neighbours"] > 250 and st["list_capacity"] >= st["max_neighbours"]


def test_pair_lists_hot_run_with_rebuilds_matches_oracle():
    """300 steps at 60 K (a rebuild every few steps, atoms change cells and slot partners): PE / KE traces vs the oracle."""
    atoms = fcc_argon(10, temperature=60.0, seed=21)
    orc = make_oracle(atoms, {(1, 1): argon_pair(8.5)})
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref = orc.run_nve(x, v, np.zeros_like(x), atoms.type_ids, 0.25, 300)
    mgr = make_manager(skin=SKIN, rc=8.5, variant=9)
    mgr.attach(atoms)
    pe0 = mgr.compute()
    assert abs(pe0 - ref[0, 0]) <= ENERGY_TOL * abs(ref[0, 0])
    th = mgr.step_nve(0.25, 300)
    assert mgr.stats()["n_builds"] >= 10
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL
    # sum r.F cancels heavily (|sum| << sum |r.F|): the pressure inherits that conditioning, hence 1e-6 on P
    p_gpu = np.array([atoms.pressure(k, w) for k, w in zip(th["ke"], th["virial_ref"])])
    assert np.max(np.abs(p_gpu - ref[1:, 4]) / np.maximum(np.abs(ref[1:, 4]), 1e-6)) <= 1e-6
    mgr.download(atoms)
    assert np.abs(atoms.positions - x).max() < 1e-8


def test_pair_lists_agree_with_the_per_atom_kernels():
    """Same decisions, same pair terms, another summation order: pair-list forces vs k_force_v3's within rounding, and the
    reassembled rows identical to the per-atom rows."""
    atoms = fcc_argon(12, temperature=50.0, seed=7, jitter=0.25)
    res = {}
    for variant in (6, 9):
        a = Atoms(atoms.type_ids, atoms.masses, atoms.positions.copy(), atoms.sim_box, velocities=atoms.velocities.copy())
        m = make_manager(skin=SKIN, variant=variant)
        m.attach(a)
        pe = m.compute()
        m.download(a, positions=False, velocities=False)
        res[variant] = (pe, a.forces.copy(), m.neighbours(a.n_atoms), m.stats())
    assert abs(res[9][0] - res[6][0]) <= 1e-13 * abs(res[6][0])
    assert force_rel_err(res[9][1], res[6][1]).max() <= 1e-12
    for r9, r6 in zip(res[9][2], res[6][2]):
        assert np.array_equal(r9, r6)
    assert res[9][3]["max_neighbours"] == res[6][3]["max_neighbours"]


def test_two_types_keep_the_per_atom_lists():
    """The pair path is single-type; a second atom type must fall back to the per-atom kernels, silently and correctly."""
    from pis_b200 import LennardJones

    table = {(1, 1): LennardJones(0.238, 3.405, 8.5), (1, 2): LennardJones(0.15, 3.0, 7.5), (2, 2): LennardJones(0.07, 2.8, 7.0)}
    atoms = fcc_argon(8, temperature=30.0, seed=9, jitter=0.1)
    atoms.type_ids[::3] = 2
    atoms.masses = [39.948, 20.18]
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    mgr = make_manager(skin=SKIN, variant=9, table=table)
    pe = mgr.compute_potential(atoms)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
