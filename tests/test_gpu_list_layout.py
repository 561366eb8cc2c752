"""List-layout options of the v3 build: the per-row x window (`build_window`) and warp-aligned rows (`list_align`).

Both are layout / work-saving choices only.  The window drops candidates the distance test would reject anyway; alignment
inserts pads (entries with the sign bit set) that every consumer skips.  So the neighbour SETS must stay bit-exact against the
oracle (lennard_jones.rs:345-415 with rcut + skin), the entry ORDER of the real entries must be the unaligned build's, and
forces / energies / traces must meet the usual bars -- with the step kernels that consume the list (thread per atom, four
lanes per atom, eight lanes per atom, the general v1 loop)."""
import numpy as np
import pytest

from pis_b200 import Atoms, LennardJones
from pis_b200.lattice import fcc_argon
from tests.helpers import SKIN, argon_pair, csr_rows_sorted, force_rel_err, make_manager, make_oracle

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-10
ENERGY_TOL = 1e-9


def _manager(align, window, variant, table):
    mgr = make_manager(skin=SKIN, variant=variant, table=table)
    mgr.set_option("list_align", align)
    mgr.set_option("build_window", window)
    return mgr


@pytest.mark.parametrize("align,window", [(0, 0), (0, 1), (1, 1), (2, 1), (2, 0)])
@pytest.mark.parametrize("variant", [7, 11, 8, 6])
def test_list_layouts_give_the_oracle_sets_and_forces(variant, align, window):
    """13^3 cells (70 A): interior warps (window + packed path) and boundary warps (minimum image) both exist."""
    atoms = fcc_argon(13, temperature=35.0, seed=21, jitter=0.3)
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    mgr = _manager(align, window, variant, table)
    mgr.attach(atoms)
    rows = mgr.neighbours(atoms.n_atoms)
    ref_rows = csr_rows_sorted(start, nbr)
    bad = [i for i in range(atoms.n_atoms) if not np.array_equal(rows[i], ref_rows[i])]
    assert not bad, f"{len(bad)} rows differ, first atom {bad[0]}"
    st = mgr.list_stats()
    assert st["listed"] == len(nbr)
    assert (st["index_words"] > st["listed"]) == (align != 0)
    pe = mgr.compute()
    mgr.download(atoms, positions=False, velocities=False)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL


def test_aligned_rows_keep_the_entry_order_and_bitwise_results():
    """Pads sit between real entries; the real entries keep their order, so the thread-per-atom kernel sums the same terms in
    the same order: forces, PE and a 60-step trace with rebuilds are BIT-identical with and without alignment."""
    base = fcc_argon(14, temperature=43.0, seed=4, jitter=0.2)
    out = []
    for align, window in [(0, 0), (1, 1), (2, 1)]:
        atoms = Atoms(base.type_ids, list(base.masses), base.positions.copy(), base.sim_box, velocities=base.velocities.copy())
        mgr = _manager(align, window, 7, {(1, 1): argon_pair()})
        mgr.attach(atoms)
        nn, nb = mgr.neighbours_padded(atoms.n_atoms)
        pe = mgr.compute()
        th = mgr.step_nve(0.25, 60)
        mgr.download(atoms)
        out.append((nn, nb, pe, th, atoms.positions.copy(), atoms.forces.copy(), mgr.stats()["n_builds"]))
    nn0, nb0, pe0, th0, x0, f0, builds0 = out[0]
    assert builds0 >= 3
    for nn, nb, pe, th, x, f, builds in out[1:]:
        assert np.array_equal(nn, nn0)
        for i in range(0, len(nn), 97):
            assert np.array_equal(nb[i, : nn[i]], nb0[i, : nn0[i]]), f"entry order of atom {i}"
        assert pe == pe0
        assert builds == builds0
        assert np.array_equal(th["pe"], th0["pe"]) and np.array_equal(th["ke"], th0["ke"])
        assert np.array_equal(x, x0) and np.array_equal(f, f0)


@pytest.mark.parametrize("align", [1, 2])
def test_aligned_rows_hot_run_follows_the_oracle(align):
    atoms = fcc_argon(12, temperature=60.0, seed=9, jitter=0.15)
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    x, v, f = atoms.positions.copy(), atoms.velocities.copy(), np.zeros_like(atoms.positions)
    ref = orc.run_nve(x, v, f, atoms.type_ids, 0.25, 120)
    mgr = _manager(align, 1, 0, table)
    mgr.attach(atoms)
    mgr.compute()
    th = mgr.step_nve(0.25, 120)
    assert mgr.stats()["n_builds"] >= 5
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL


@pytest.mark.parametrize("align", [0, 2])
def test_window_and_alignment_with_per_pair_cutoffs(align):
    """Three types, different list radii per pair, one pair missing: the window must use the LARGEST list radius."""
    atoms = fcc_argon(12, temperature=30.0, seed=5, jitter=0.25)
    n = atoms.n_atoms
    types = (1 + (np.arange(n) * 7 + (np.arange(n) // 5)) % 3).astype(np.int32)
    atoms = Atoms(types, [39.948, 20.18, 83.798], atoms.positions, atoms.sim_box, velocities=atoms.velocities)
    table = {(1, 1): LennardJones(0.238, 3.405, 8.5), (1, 2): LennardJones(0.15, 3.0, 7.5), (2, 2): LennardJones(0.07, 2.8, 7.0),
             (1, 3): LennardJones(0.3, 3.5, 9.0), (3, 3): LennardJones(0.4, 3.6, 6.2)}
    orc = make_oracle(atoms, table)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    mgr = _manager(align, 1, 7, table)
    mgr.attach(atoms)
    pe = mgr.compute()
    rows = mgr.neighbours(n)
    ref_rows = csr_rows_sorted(start, nbr)
    bad = [i for i in range(n) if not np.array_equal(rows[i], ref_rows[i])]
    assert not bad, f"{len(bad)} rows differ, first atom {bad[0]}"
    mgr.download(atoms, positions=False, velocities=False)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL


def test_window_on_a_non_cubic_box_with_vacuum():
    """Cell edges differ per dimension and half the box is empty: the window works from each dimension's own cell edge."""
    base = fcc_argon(12, temperature=25.0, seed=8, jitter=0.1)
    a = 5.41
    keep = base.positions[:, 2] < 6 * a
    pos = base.positions[keep][:-1]
    from pis_b200 import SimulationBox
    box = SimulationBox.from_lammps_data(0.0, 12 * a, 0.0, 12 * a * 1.13, 0.0, 12 * a * 1.31)
    atoms = Atoms(np.ones(len(pos), dtype=np.int32), [39.948], pos * np.array([1.0, 1.13, 1.0]), box,
                  velocities=base.velocities[keep][:-1])
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN)
    for align in (0, 1):
        mgr = _manager(align, 1, 7, table)
        mgr.attach(atoms)
        rows = mgr.neighbours(atoms.n_atoms)
        ref_rows = csr_rows_sorted(start, nbr)
        bad = [i for i in range(atoms.n_atoms) if not np.array_equal(rows[i], ref_rows[i])]
        assert not bad, f"align {align}: {len(bad)} rows differ, first atom {bad[0]}"
