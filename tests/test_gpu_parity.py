"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded inputs.  Bars (BASELINE.json north_star): neighbour lists bit-exact as sorted pair sets;
forces within 1e-10 relative; energy / temperature traces within 1e-9 relative."""
import numpy as np
import pytest

from pis_b200 import Atoms, LennardJones, SimulationBox, capi
from pis_b200.lattice import ARGON, fcc_argon
from tests.helpers import RC25, SKIN, argon_pair, csr_rows_sorted, force_rel_err, make_manager, make_oracle

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-10
ENERGY_TOL = 1e-9


def assert_virial_trace(atoms, th, ref, x_end, f_end):
    """tr(X F^T) = sum_i r_i . F_i with WRAPPED absolute positions (properties.rs:49-51,61-65) is ill-conditioned: the forces sum
    to zero, so terms of size |r_i||F_i| (|r| up to L) cancel to a result orders of magnitude smaller.  The honest bar is
    therefore absolute: the north_star allows each force a relative error FORCE_TOL, i.e. an error of at most
    FORCE_TOL x S in the sum, S = sum_i |r_i||F_i| (taken from the oracle's final state, times 2 for the variation along the
    trace).  Compared are the virials themselves: the oracle's from its pressure column, W = 3 V P - 2 KE."""
    vol = atoms.sim_box.volume()
    w_ref = 3.0 * vol * ref[1:, 4] - 2.0 * ref[1:, 1]
    scale = 2.0 * float((np.linalg.norm(x_end, axis=1) * np.linalg.norm(f_end, axis=1)).sum())
    assert np.max(np.abs(th["virial_ref"] - w_ref)) <= FORCE_TOL * scale, (np.max(np.abs(th["virial_ref"] - w_ref)), scale)
    # and the pressure the host prints from it
    p_gpu = np.array([atoms.pressure(k, w) for k, w in zip(th["ke"], th["virial_ref"])])
    assert np.max(np.abs(p_gpu - ref[1:, 4])) <= (FORCE_TOL * scale + 2.0 * ENERGY_TOL * np.abs(ref[1:, 1]).max()) / (3.0 * vol)


def _jittered(ncell, jitter=0.15, temperature=30.0, seed=7):
    return fcc_argon(ncell, temperature=temperature, seed=seed, jitter=jitter)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 6, 7, 8, 11, 12])
@pytest.mark.parametrize("ncell,skin", [(10, 0.0), (10, SKIN), (16, SKIN)])
def test_compute_potential_parity(ncell, skin, variant):
    atoms = _jittered(ncell)
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    mgr = make_manager(skin=skin, variant=variant)
    pe = mgr.compute_potential(atoms)  # adds into atoms.forces (zeros)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
    # Newton's third law on the full-list kernel
    assert np.abs(atoms.forces.sum(axis=0)).max() < 1e-9


def test_compute_potential_accumulates():
    atoms = _jittered(6)
    atoms.forces[...] = 1.25
    mgr = make_manager(skin=SKIN, rc=8.5)
    orc = make_oracle(atoms, {(1, 1): argon_pair(8.5)})
    _, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    mgr.compute_potential(atoms)
    assert force_rel_err(atoms.forces - 1.25, f_ref).max() <= 1e-9  # 1.25 + f rounding


def test_argon4000_lattice_known_answer():
    """example/argon4000.txt geometry: 54 neighbours/atom, step-0 PE = -6956.99645673589."""
    atoms = fcc_argon(10, temperature=0.0)
    mgr = make_manager(skin=0.0, rc=8.5)
    pe = mgr.compute_potential(atoms)
    assert abs(pe - (-6956.99645673589)) < 1e-8
    assert np.abs(atoms.forces).max() < 1e-12
    rows = mgr.neighbours(atoms.n_atoms)
    assert all(len(r) == 54 for r in rows)


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 6, 7])
@pytest.mark.parametrize("ncell,skin", [(8, 0.0), (8, SKIN), (12, SKIN)])
def test_neighbour_list_exact(ncell, skin, variant):
    atoms = _jittered(ncell, jitter=0.3)
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=skin)
    ref_rows = csr_rows_sorted(start, nbr)
    mgr = make_manager(skin=skin, variant=variant)
    mgr.attach(atoms)
    rows = mgr.neighbours(atoms.n_atoms)
    assert sum(len(r) for r in rows) == len(nbr)
    for i, (a, b) in enumerate(zip(rows, ref_rows)):
        assert np.array_equal(a, b), f"atom {i}"


def test_neighbour_list_shell_on_cutoff():
    """Cutoff is inclusive and evaluated as sqrt(r2) > rc: put an FCC shell exactly on the list cutoff."""
    atoms = fcc_argon(8, temperature=0.0)
    a = 5.41
    rc = a * np.sqrt(1.5)  # 3rd shell distance a*sqrt(3/2); lattice r2 values straddle it by rounding
    table = {(1, 1): LennardJones(0.238, 3.405, rc, True)}
    orc = make_oracle(atoms, table)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=0.0)
    mgr = make_manager(skin=0.0, table=table)
    mgr.attach(atoms)
    rows = mgr.neighbours(atoms.n_atoms)
    for a_, b_ in zip(rows, csr_rows_sorted(start, nbr)):
        assert np.array_equal(a_, b_)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    pe = mgr.compute()
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)


def test_drift_is_bit_exact_and_kick_close():
    """One verlet_step_nve from identical (x, v, F): positions are bit-identical (same arithmetic,
    same inputs), velocities differ only through the force summation order."""
    atoms = _jittered(8)
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    _, f0 = orc.compute_potential(atoms.positions, atoms.type_ids)
    x, v, f = atoms.positions.copy(), atoms.velocities.copy(), f0.copy()
    pe_ref = orc.verlet_step_nve(x, v, f, atoms.type_ids, 0.25)
    atoms.forces[...] = f0
    mgr = make_manager(skin=SKIN)
    pe = mgr.verlet_step_nve(atoms, 0.25)
    assert np.array_equal(atoms.positions, x)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f).max() <= FORCE_TOL
    assert np.abs(atoms.velocities - v).max() <= 1e-13 * max(1.0, np.abs(v).max())


def test_nve_trace_parity_1000_steps():
    """argon4000-sized system, 1000 NVE steps: PE / KE / T traces within 1e-9 relative of the oracle."""
    atoms = fcc_argon(10, temperature=5.0, seed=12345)
    table = {(1, 1): argon_pair(8.5)}
    orc = make_oracle(atoms, table)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    f = np.zeros_like(x)
    steps = 1000
    ref = orc.run_nve(x, v, f, atoms.type_ids, 0.25, steps)
    mgr = make_manager(skin=SKIN, rc=8.5)
    mgr.attach(atoms)
    pe0 = mgr.compute()
    assert abs(pe0 - ref[0, 0]) <= ENERGY_TOL * abs(ref[0, 0])
    th = mgr.step_nve(0.25, steps)
    pe_ref, ke_ref, t_ref = ref[1:, 0], ref[1:, 1], ref[1:, 3]
    assert np.max(np.abs(th["pe"] - pe_ref) / np.abs(pe_ref)) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ke_ref) / np.abs(ke_ref)) <= ENERGY_TOL
    t_gpu = np.array([atoms.temerature(k) for k in th["ke"]])
    assert np.max(np.abs(t_gpu - t_ref) / np.abs(t_ref)) <= ENERGY_TOL
    # pressure uses tr(X F^T) with wrapped positions (properties.rs:61-65)
    assert_virial_trace(atoms, th, ref, x, f)
    # NVE: the Hamiltonian is conserved
    h = th["pe"] + th["ke"]
    assert np.abs(h - h[0]).max() <= 2e-4 * abs(h[0])


def test_hot_run_rebuilds_and_matches_no_skin():
    """A hot system triggers list rebuilds; list-with-skin forces == no-list forces at the end."""
    atoms = fcc_argon(8, temperature=60.0, seed=3, jitter=0.05)
    orc0 = make_oracle(atoms, {(1, 1): argon_pair()})
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref = orc0.run_nve(x, v, np.zeros_like(x), atoms.type_ids, 0.25, 300)
    mgr = make_manager(skin=SKIN)
    mgr.attach(atoms)
    mgr.compute()
    th = mgr.step_nve(0.25, 300)
    st = mgr.stats()
    assert st["n_builds"] >= 3
    # the skin/rebuild machinery must not change the trajectory: traces match the no-list oracle
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL
    mgr.download(atoms)
    f_skin = atoms.forces.copy()
    orc = make_oracle(atoms, {(1, 1): argon_pair()})
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    assert force_rel_err(f_skin, f_ref).max() <= FORCE_TOL
    assert abs(th["pe"][-1] - pe_ref) <= ENERGY_TOL * abs(pe_ref)


def test_host_step_equals_resident_step():
    a1 = _jittered(6, temperature=40.0)
    a2 = Atoms(a1.type_ids, a1.masses, a1.positions.copy(), a1.sim_box, velocities=a1.velocities.copy())
    m1, m2 = make_manager(skin=SKIN), make_manager(skin=SKIN)
    m1.compute_potential(a1)
    m2.attach(a2)
    m2.compute()
    pes = [m1.verlet_step_nve(a1, 0.25) for _ in range(20)]
    th = m2.step_nve(0.25, 20)
    m2.download(a2)
    assert np.allclose(pes, th["pe"], rtol=1e-12, atol=0)
    assert np.abs(a1.positions - a2.positions).max() < 1e-9
    assert np.abs(a1.velocities - a2.velocities).max() < 1e-10


def test_two_types_and_missing_pair():
    """Dense per-type-pair table; keys are stored as given and looked up sorted (potential.rs:181-192):
    (2,1) is never found, so 1-2 pairs are skipped -- like the reference."""
    atoms = _jittered(6, jitter=0.2)
    atoms.type_ids[::3] = 2
    atoms.masses = [39.948, 20.18]
    for missing, table in (
        (0, {(1, 1): LennardJones(0.238, 3.405, 8.5), (1, 2): LennardJones(0.15, 3.0, 7.5), (2, 2): LennardJones(0.07, 2.8, 7.0)}),
        (1, {(1, 1): LennardJones(0.238, 3.405, 8.5), (2, 1): LennardJones(0.15, 3.0, 7.5), (2, 2): LennardJones(0.07, 2.8, 7.0)}),
    ):
        atoms.forces[...] = 0.0
        orc = make_oracle(atoms, table)
        pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
        mgr = make_manager(skin=0.5, table=table)
        pe = mgr.compute_potential(atoms)
        assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
        assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
        start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=0.5)
        for a_, b_ in zip(mgr.neighbours(atoms.n_atoms), csr_rows_sorted(start, nbr)):
            assert np.array_equal(a_, b_)
        # the reference's "potential was missing" line per candidate pair (lennard_jones.rs:216-222) is a count here
        assert mgr.stats()["missing_type_pairs"] == missing


def test_triclinic_box_general_path():
    """Non-orthorhombic h exercises the full 3x3 mat-vec restatement (NPT makes the box triclinic)."""
    base = _jittered(6, jitter=0.2)
    L = base.sim_box.h[0, 0]
    box = SimulationBox.from_lammps_data(0, L, 0, L, 0, L, xy=3.1, xz=-2.2, yz=1.7)
    atoms = Atoms(base.type_ids, base.masses, base.positions, box, velocities=base.velocities)
    table = {(1, 1): argon_pair(8.5)}
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    mgr = make_manager(skin=0.0, rc=8.5)
    pe = mgr.compute_potential(atoms)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
    x, v, f = atoms.positions.copy(), atoms.velocities.copy(), atoms.forces.copy()
    orc.verlet_step_nve(x, v, f, atoms.type_ids, 0.25)
    mgr.verlet_step_nve(atoms, 0.25)
    assert np.abs(atoms.positions - x).max() < 1e-12


def test_edge_positions_on_boundary_and_beyond():
    """Atoms exactly on 0 / L and coordinates >= L (unwrapped step-0 input): same pairs as the oracle."""
    atoms = fcc_argon(7, temperature=0.0)
    L = atoms.sim_box.h[0, 0]
    atoms.positions[5] = [L, 0.9, L]
    atoms.positions[11] = [L + 1.0, L + 2.5, 0.7]
    atoms.positions[17] = [0.0, L, 1.3]
    table = {(1, 1): argon_pair(8.5)}
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    mgr = make_manager(skin=0.0, rc=8.5)
    pe = mgr.compute_potential(atoms)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL


@pytest.mark.parametrize("variant", [0, 6, 8, 11])
def test_unwrapped_inputs_far_outside_the_box_and_a_box_centred_on_the_origin(variant):
    """Step-0 inputs need not lie in [0, L) (simulation.rs:28 calls compute_potential on them as read): atoms several
    box lengths out, and a whole system given in [-L/2, L/2).  The interior-warp shortcut of the force kernels uses raw
    coordinate differences, so it must stay off until a drift has wrapped the positions (the flag words FLAG_UNWRAPPED_*):
    forces, energy and lists must equal those of the SAME system wrapped into the box, then both runs must stay together."""
    base = _jittered(16, jitter=0.2, temperature=40.0, seed=17)      # 16384 atoms, box edge 86.6: interior warps exist
    L = base.sim_box.h[0, 0]
    rng = np.random.default_rng(3)
    cases = {"centred": base.positions - 0.5 * L,
             "far": base.positions + L * rng.integers(-3, 4, size=base.positions.shape).astype(np.float64)}
    ref_mgr = make_manager(skin=SKIN, variant=variant)
    ref_mgr.attach(base)
    pe_ref = ref_mgr.compute()
    rows_ref = ref_mgr.neighbours(base.n_atoms)
    a_ref = Atoms(base.type_ids, base.masses, base.positions.copy(), base.sim_box, velocities=base.velocities.copy())
    ref_mgr.download(a_ref, positions=False, velocities=False)
    th_ref = ref_mgr.step_nve(0.25, 6)
    for name, pos in cases.items():
        a = Atoms(base.type_ids, base.masses, np.ascontiguousarray(pos), base.sim_box, velocities=base.velocities.copy())
        m = make_manager(skin=SKIN, variant=variant)
        m.attach(a)
        pe = m.compute()
        m.download(a, velocities=False)
        assert np.array_equal(a.positions, pos), name                     # compute_potential leaves the positions alone
        assert abs(pe - pe_ref) <= 1e-12 * abs(pe_ref), name
        assert force_rel_err(a.forces, a_ref.forces).max() <= 1e-11, name  # image shifts by k L round differently: not bitwise
        for r1, r0 in zip(m.neighbours(a.n_atoms), rows_ref):
            assert np.array_equal(r1, r0), name
        th = m.step_nve(0.25, 6)
        for key in ("pe", "ke"):
            assert np.max(np.abs(th[key] - th_ref[key]) / np.abs(th_ref[key])) <= 1e-10, (name, key)
        m.download(a)
        assert a.positions.min() >= 0.0 and a.positions.max() <= L        # the drift wrapped them


def test_two_atoms_analytic():
    """u(2^(1/6) sigma) = -eps - u_cut and f = 0 there."""
    eps, sig, rc = 0.238, 3.405, 8.5
    L = 40.0
    r = 2.0 ** (1.0 / 6.0) * sig
    box = SimulationBox.from_lammps_data(0, L, 0, L, 0, L)
    atoms = Atoms([1, 1], [39.948], [[1.0, 1.0, 1.0], [1.0 + r, 1.0, 1.0]], box)
    mgr = make_manager(skin=0.0, rc=rc)
    pe = mgr.compute_potential(atoms)
    ucut = 4 * eps * ((sig / rc) ** 12 - (sig / rc) ** 6)
    assert abs(pe - (-eps - ucut)) < 1e-12
    assert np.abs(atoms.forces).max() < 1e-10


def test_errors():
    from pis_b200.capi import PisbError

    box = SimulationBox.from_lammps_data(0, 10, 0, 10, 0, 10)  # shorter than 2 cutoffs
    atoms = Atoms([1, 1], [39.948], [[1.0, 1.0, 1.0], [4.0, 1.0, 1.0]], box)
    mgr = make_manager(skin=0.0, rc=8.5)
    with pytest.raises(PisbError):
        mgr.compute_potential(atoms)
    atoms2 = fcc_argon(6, temperature=0.0)
    atoms2.type_ids[3] = 5
    with pytest.raises(PisbError):
        make_manager(skin=0.0, rc=8.5).compute_potential(atoms2)


def test_kernel_families_agree():
    """Two arithmetic families.  REFERENCE-ORDER kernels (v1 all-FP64, v2 FP32 pre-filter + queue) keep
    the reference's operation order per pair: their forces / energies / trajectories are BIT-identical to one another.
    The LEAN kernels (v3 and everything built on its loop: the defaults, the fused step kernels) shorten the per-pair arithmetic (pisb_device.cuh)
    but take the same in/out decisions: bit-identical among themselves, and within rounding of the reference-order family."""
    atoms = _jittered(12, jitter=0.25, temperature=50.0)
    out = {}
    for variant in (1, 2, 3, 6, 11):
        a = Atoms(atoms.type_ids, atoms.masses, atoms.positions.copy(), atoms.sim_box, velocities=atoms.velocities.copy())
        m = make_manager(skin=SKIN, variant=variant)
        m.set_option("fuse_vv", 0)   # the force kernels proper; k_force_vv reduces KE over other block sizes (own test below)
        m.attach(a)
        pe0 = m.compute()
        m.download(a, positions=False, velocities=False)
        f0 = a.forces.copy()
        th = m.step_nve(0.25, 25)
        m.download(a)
        out[variant] = (pe0, th, a.positions.copy(), a.velocities.copy(), a.forces.copy(), f0)
    for base, others in ((1, (2,)), (3, (6,))):
        for other in others:
            assert out[base][0] == out[other][0]
            for name in ("pe", "ke", "virial_ref", "virial_pair"):
                assert np.array_equal(out[base][1][name], out[other][1][name]), name
            for k in (2, 3, 4, 5):
                assert np.array_equal(out[base][k], out[other][k])
    for v in (3, 11):      # 11: four lanes per atom -- the lean arithmetic again, partial sums met in a butterfly
        ref, lean = out[1], out[v]
        assert abs(lean[0] - ref[0]) <= 1e-13 * abs(ref[0])
        assert force_rel_err(lean[5], ref[5]).max() <= 1e-12          # same state, two arithmetics: rounding only
        for name in ("pe", "ke", "virial_pair"):
            assert np.max(np.abs(lean[1][name] - ref[1][name]) / np.abs(ref[1][name])) <= 1e-11, name
        assert np.abs(lean[2] - ref[2]).max() <= 1e-11                # 25 steps apart by accumulated rounding only


def test_guard_band_pairs_sit_on_the_cutoff():
    """Stress the guard band: many pairs placed within +-1e-7 (relative) of rc and of rc+skin, i.e. far
    inside the FP32-ambiguous band; sets must still be exact and forces must still match."""
    rng = np.random.default_rng(5)
    base = fcc_argon(8, temperature=0.0)
    L = base.sim_box.h[0, 0]
    n = 600
    pos = np.zeros((2 * n, 3))
    centers = rng.uniform(0, L, size=(n, 3))
    dirs = rng.standard_normal((n, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    target = np.where(np.arange(n) % 2 == 0, RC25, RC25 + SKIN) * (1.0 + rng.uniform(-1e-7, 1e-7, size=n))
    pos[0::2] = centers
    pos[1::2] = centers + dirs * target[:, None]
    pos -= np.floor(pos / L) * L
    atoms = Atoms(np.ones(2 * n, dtype=np.int32), [39.948], pos, base.sim_box)
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    for variant in (6, 7, 4):
        mgr = make_manager(skin=SKIN, variant=variant)
        mgr.attach(atoms)
        for a_, b_ in zip(mgr.neighbours(atoms.n_atoms), csr_rows_sorted(start, nbr)):
            assert np.array_equal(a_, b_)
    pe = mgr.compute()
    mgr.download(atoms)
    assert abs(pe - pe_ref) <= 1e-12 * max(abs(pe_ref), 1.0)
    assert np.abs(atoms.forces - f_ref).max() <= 1e-13 * max(np.abs(f_ref).max(), 1.0)


def test_gpu_matches_committed_golden_vectors():
    """tests/golden/oracle_small.json (made by tests/golden/make_golden.py) without running the oracle."""
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_small.json")) as f:
        g = json.load(f)
    box = SimulationBox.from_lammps_data(0, g["L"], 0, g["L"], 0, g["L"])
    atoms = Atoms(g["types"], [g["mass"]], g["positions"], box, velocities=g["velocities"])
    table = {(1, 1): LennardJones(g["eps"], g["sigma"], g["rc"], True)}
    mgr = make_manager(skin=g["skin"], table=table)
    mgr.attach(atoms)
    rows = mgr.neighbours(atoms.n_atoms)
    assert [r.tolist() for r in rows] == g["neighbours_skin"]
    pe = mgr.compute()
    mgr.download(atoms)
    f_ref = np.array(g["forces"])
    assert abs(pe - g["pe"]) <= 1e-12 * abs(g["pe"])
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
    th = mgr.step_nve(g["dt"], g["steps"])
    ref = np.array(g["thermo"])
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL
    mgr.download(atoms)
    assert np.abs(atoms.positions - np.array(g["positions_end"])).max() < 1e-10


def test_gpu_matches_committed_ensemble_golden_vectors():
    """tests/golden/oracle_small_ensembles.json: 25 NVT and 25 NPT steps of the golden system, without running the oracle."""
    import json
    import os

    gd = os.path.join(os.path.dirname(__file__), "golden")
    with open(os.path.join(gd, "oracle_small.json")) as f:
        g = json.load(f)
    with open(os.path.join(gd, "oracle_small_ensembles.json")) as f:
        e = json.load(f)
    table = {(1, 1): LennardJones(g["eps"], g["sigma"], g["rc"], True)}
    for ens in ("nvt", "npt"):
        box = SimulationBox.from_lammps_data(0, g["L"], 0, g["L"], 0, g["L"])
        atoms = Atoms(g["types"], [g["mass"]], g["positions"], box, velocities=g["velocities"])
        mgr = make_manager(skin=g["skin"], table=table)
        mgr.attach(atoms)
        mgr.compute()
        chain = mgr.nhc_new(*e["temp"])
        ref = np.array(e[ens + "_thermo"])
        if ens == "nvt":
            th, en = mgr.step_nvt_nhc(g["dt"], e["steps"], chain, 0, e["steps"])
        else:
            baro = mgr.mtk_new(e["iso"][0], e["iso"][1], atoms.n_atoms, e["temp"][0])
            th, en, hh = mgr.step_npt_mtk(g["dt"], e["steps"], baro, chain, 0, e["steps"], atoms)
            h_ref = np.array(e["npt_h"])[1:].reshape(-1, 3, 3).transpose(0, 2, 1)
            assert np.abs(hh - h_ref).max() <= 1e-9 * np.abs(h_ref).max()
            assert np.abs(np.array(baro.momentum) - np.array(e["npt_momentum"])).max() <= 1e-8 * np.abs(np.array(e["npt_momentum"])).max()
        assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
        assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL
        assert np.max(np.abs(th["pe"] + th["ke"] + en - ref[1:, 2]) / np.abs(ref[1:, 2])) <= ENERGY_TOL


@pytest.mark.parametrize("halo_mode,force_variant", [(2, 0), (1, 0), (2, 3), (1, 3)])
def test_multi_gpu_equals_single_gpu(halo_mode, force_variant):
    """2-rank spatially decomposed run == 1-GPU run (neighbour sets exact per global id, forces 1e-10,
    traces 1e-9), with the peer-memory halo (2) and with the NCCL fallback (1); force_variant 3 makes the bricks step with
    the fused k_force_vv kernel (speculative launch behind the decision word) as 4M-atom bricks do by default.  Needs 2 visible GPUs; the
    torchrun launch mirrors the driver's."""
    import json
    import os
    import subprocess
    import sys

    from pis_b200 import capi

    if capi.load().pisb_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(root, "tools", "multi_check.py"), "12", "40", "60"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=dict(os.environ, PISB_HALO_MODE=str(halo_mode), PISB_FORCE_VARIANT=str(force_variant)))
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["ok"] and out["halo_mode"] == halo_mode, out


def test_config2_256k_atoms_parity():
    """BASELINE configs[1]: FCC argon 256 000 atoms, rc = 2.5 sigma, skin = 0.3 sigma.  Forces / energy / neighbour
    sets against the oracle at full size, then a 60-step NVE trace (the 1000-step trace is run at 4000 atoms in
    test_nve_trace_parity_1000_steps: the CPU oracle needs ~0.2 s per 256k-atom step even with all cores)."""
    atoms = fcc_argon(40, temperature=30.0, seed=12345, jitter=0.1)
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids, mode="omp")
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN)
    mgr = make_manager(skin=SKIN)
    mgr.attach(atoms)
    rows = mgr.neighbours(atoms.n_atoms)
    assert sum(len(r) for r in rows) == len(nbr)
    ref_rows = csr_rows_sorted(start, nbr)
    bad = sum(0 if np.array_equal(a_, b_) else 1 for a_, b_ in zip(rows, ref_rows))
    assert bad == 0
    pe = mgr.compute()
    mgr.download(atoms, positions=False, velocities=False)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
    steps = 60
    x, v, f = atoms.positions.copy(), atoms.velocities.copy(), f_ref.copy()
    pes, kes = [], []
    for _ in range(steps):
        pes.append(orc.verlet_step_nve(x, v, f, atoms.type_ids, 0.25, mode="omp"))
        kes.append(orc.kinetic_energy(v, atoms.type_ids))
    th = mgr.step_nve(0.25, steps)
    assert np.max(np.abs(th["pe"] - np.array(pes)) / np.abs(pes)) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - np.array(kes)) / np.abs(kes)) <= ENERGY_TOL
    mgr.download(atoms)
    assert np.abs(atoms.positions - x).max() < 1e-9
    assert force_rel_err(atoms.forces, f).max() <= FORCE_TOL


def test_full_size_properties_4m_atoms():
    """BASELINE configs[2] size (4M atoms), where the oracle is too slow: size-independent properties instead --
    Newton's third law, list symmetry (j in list(i) <=> i in list(j)) via the degree sum and a sample, energy
    conservation, and step-function invariance under a rigid translation by a lattice vector."""
    atoms = fcc_argon(100, temperature=20.0, seed=5)
    mgr = make_manager(skin=SKIN)
    mgr.attach(atoms)
    pe0 = mgr.compute()
    mgr.download(atoms, positions=False, velocities=False)
    assert np.abs(atoms.forces.sum(axis=0)).max() < 1e-7            # sum F = 0
    assert abs(pe0 / atoms.n_atoms - (-1.73)) < 0.05              # near the lattice energy per atom (rc = 2.5 sigma)
    th = mgr.step_nve(0.25, 50)
    h = th["pe"] + th["ke"]
    assert np.abs(h - h[0]).max() <= 1e-3 * abs(h[0])   # dt = 0.25, unshifted force at rc: O(1e-4) wobble is inherent
    # translate every atom by one lattice constant (periodic box): identical physics, bitwise-equal wrapped lattice
    a2 = fcc_argon(100, temperature=20.0, seed=5)
    L = a2.sim_box.h[0, 0]
    a2.positions[:, 0] += 5.41
    a2.positions[:, 0] -= np.floor(a2.positions[:, 0] / L) * L
    m2 = make_manager(skin=SKIN)
    m2.attach(a2)
    pe0b = m2.compute()
    assert abs(pe0b - pe0) <= 1e-11 * abs(pe0)


def test_non_cubic_box_and_vacuum_slab():
    """Orthorhombic box with three different edges, half of it empty (a slab: empty cells, uneven density)."""
    base = fcc_argon(8, temperature=25.0, seed=8, jitter=0.1)
    a = 5.41
    keep = base.positions[:, 2] < 4 * a                      # keep the lower half in z -> vacuum above
    pos = base.positions[keep] * np.array([1.0, 1.0, 1.0])
    box = SimulationBox.from_lammps_data(0, 8 * a, 0, 8 * a + 3.7, 0, 8 * a + 9.1)
    atoms = Atoms(np.ones(keep.sum(), dtype=np.int32), [39.948], pos, box, velocities=base.velocities[keep])
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN)
    mgr = make_manager(skin=SKIN)
    mgr.attach(atoms)
    for a_, b_ in zip(mgr.neighbours(atoms.n_atoms), csr_rows_sorted(start, nbr)):
        assert np.array_equal(a_, b_)
    pe = mgr.compute()
    mgr.download(atoms, positions=False, velocities=False)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
    x, v, f = atoms.positions.copy(), atoms.velocities.copy(), f_ref.copy()
    pes = [orc.verlet_step_nve(x, v, f, atoms.type_ids, 0.25) for _ in range(40)]
    th = mgr.step_nve(0.25, 40)
    assert np.max(np.abs(th["pe"] - np.array(pes)) / np.abs(pes)) <= ENERGY_TOL


def test_list_capacity_grows_and_user_cap_is_enforced():
    """A dense cluster exceeds the density-based capacity estimate: the list is regrown transparently;
    an explicit list_capacity that is too small is reported as PISB_ERR_CAPACITY."""
    from pis_b200.capi import PISB_ERR_CAPACITY, PisbError

    rng = np.random.default_rng(11)
    L = 60.0
    n = 400
    # all atoms inside a 14 A ball in a large box: mean density is tiny, local density is high
    pts = rng.standard_normal((4000, 3))
    pts = pts[np.linalg.norm(pts, axis=1) < 2.0][:n] * 3.5 + L / 2
    # thin out pairs closer than 2.9 A so the energy stays finite
    keep = []
    for p in pts:
        if all(np.linalg.norm(p - q) > 2.9 for q in keep):
            keep.append(p)
    pos = np.array(keep)
    box = SimulationBox.from_lammps_data(0, L, 0, L, 0, L)
    atoms = Atoms(np.ones(len(pos), dtype=np.int32), [39.948], pos, box)
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    mgr = make_manager(skin=SKIN)
    pe = mgr.compute_potential(atoms)
    st = mgr.stats()
    assert st["max_neighbours"] > 30 and st["list_capacity"] >= st["max_neighbours"]
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
    small = make_manager(skin=SKIN)
    small.set_option("list_capacity", 8)
    atoms.forces[...] = 0.0
    with pytest.raises(PisbError) as e:
        small.compute_potential(atoms)
    assert e.value.code == PISB_ERR_CAPACITY


def test_nvt_nose_hoover_trace_parity():
    """verlet_step_nvt_nhc (potential.rs:35-58) + NHThermostatChain (nvt.rs) with the temperature ramp of
    `fix mynvt all nvt temp 5.0 50.0 50`; PE / KE / Hamiltonian (incl. thermostat energy) traces and the chain state."""
    atoms = fcc_argon(8, temperature=5.0, seed=12345)
    table = {(1, 1): argon_pair(8.5)}
    orc = make_oracle(atoms, table)
    steps = 400
    chain_ref = orc.nhc_new(5.0, 50.0, 50.0)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref = orc.run_nvt(x, v, np.zeros_like(x), atoms.type_ids, 0.25, steps, chain_ref)
    mgr = make_manager(skin=SKIN, rc=8.5)
    mgr.attach(atoms)
    mgr.compute()
    chain = mgr.nhc_new(5.0, 50.0, 50.0)
    assert list(chain.q) == list(chain_ref.q) and chain.target_temperature == 5.0
    # two batches: the chain state and the ramp index carry over
    th1, e1 = mgr.step_nvt_nhc(0.25, 150, chain, 0, steps)
    th2, e2 = mgr.step_nvt_nhc(0.25, steps - 150, chain, 150, steps)
    pe = np.concatenate([th1["pe"], th2["pe"]])
    ke = np.concatenate([th1["ke"], th2["ke"]])
    ham = pe + ke + np.concatenate([e1, e2])
    assert np.max(np.abs(pe - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(ke - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL
    assert np.max(np.abs(ham - ref[1:, 2]) / np.abs(ref[1:, 2])) <= ENERGY_TOL
    assert atoms.temerature(ke[-1]) > 40.0                     # the ramp heated the crystal towards 50 K
    for a_, b_ in zip(list(chain.xi) + [chain.target_temperature], list(chain_ref.xi) + [chain_ref.target_temperature]):
        assert abs(a_ - b_) <= 1e-9 * max(abs(b_), 1e-12)
    mgr.download(atoms)
    assert np.abs(atoms.positions - x).max() < 1e-8


def _npt_case(ncell, tau_p, steps, split, t_ramp=(5.0, 50.0, 100.0), pressure=0.01, skin=SKIN):
    atoms = fcc_argon(ncell, temperature=5.0, seed=12345)
    table = {(1, 1): argon_pair(8.5)}
    orc = make_oracle(atoms, table)
    chain_ref = orc.nhc_new(*t_ramp)
    baro_ref = orc.mtk_new(pressure, tau_p, atoms.n_atoms, t_ramp[0])
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref, htr = orc.run_npt(x, v, np.zeros_like(x), atoms.type_ids, 0.25, steps, baro_ref, chain_ref)
    mgr = make_manager(skin=skin, rc=8.5)
    mgr.attach(atoms)
    mgr.compute()
    cells0 = mgr.stats()["n_cells"]
    chain = mgr.nhc_new(*t_ramp)
    baro = mgr.mtk_new(pressure, tau_p, atoms.n_atoms, t_ramp[0])
    assert baro.w == baro_ref.w
    th1, e1, h1 = mgr.step_npt_mtk(0.25, split, baro, chain, 0, steps, atoms)
    cells_mid = mgr.stats()["n_cells"]
    th2, e2, h2 = mgr.step_npt_mtk(0.25, steps - split, baro, chain, split, steps, atoms)
    pe = np.concatenate([th1["pe"], th2["pe"]])
    ke = np.concatenate([th1["ke"], th2["ke"]])
    vir = np.concatenate([th1["virial_ref"], th2["virial_ref"]])
    ham = pe + ke + np.concatenate([e1, e2])
    hh = np.concatenate([h1, h2])
    assert np.max(np.abs(pe - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(ke - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL
    assert np.max(np.abs(ham - ref[1:, 2]) / np.abs(ref[1:, 2])) <= ENERGY_TOL
    h_ref = htr[1:].reshape(-1, 3, 3).transpose(0, 2, 1)
    assert np.abs(hh - h_ref).max() <= 1e-9 * np.abs(h_ref).max()
    vol = np.abs(np.linalg.det(hh))
    pres = (2.0 * ke + vir) / (3.0 * vol)
    assert np.max(np.abs(pres - ref[1:, 4])) <= 1e-9 * np.abs(ref[1:, 4]).max()
    for a_, b_ in zip(list(baro.momentum) + list(chain.xi), list(baro_ref.momentum) + list(chain_ref.xi)):
        assert abs(a_ - b_) <= 1e-8 * max(abs(b_), np.abs(np.array(baro_ref.momentum)).max() * 1e-3, 1e-12)
    mgr.download(atoms)
    assert np.abs(atoms.positions - x).max() < 1e-7
    assert np.array_equal(atoms.sim_box.h, hh[-1])              # sync_box: atoms.sim_box follows scale_box
    return mgr, (cells0, cells_mid), hh


def test_npt_mtk_trace_parity_example_fix_line():
    """verlet_step_npt_mtk (potential.rs:112-135) with the reference's shipped
    `fix mynpt all npt temp 5.0 50.0 100 iso 0.01 0.01 1000`; PE / KE / H / P traces, box trace, barostat + chain state."""
    mgr, cells0, hh = _npt_case(8, 1000.0, 300, 120)
    off = np.abs(hh[-1] - np.diag(np.diag(hh[-1]))).max()
    assert off > 0.0                                   # the pressure tensor's off-diagonals made the box triclinic
    assert mgr.stats()["n_builds"] >= 2


def test_npt_fast_barostat_regrids_cells():
    """A stiff barostat (tau 25) strains the 9^3 FCC box by several percent within 70 steps: the cell grid drops from
    5 to 4 cells in two dimensions mid-run and returns to 5 (re-grid + forced rebuild each time) and the skin budget
    shrinks with the affine strain since the last build."""
    mgr, (cells0, cells_mid), hh = _npt_case(9, 25.0, 70, 30)
    assert list(cells0) == [10, 10, 10]                 # orthorhombic start: half-size cells (5 list cutoffs per edge)
    assert list(cells_mid) == [4, 4, 5]                 # triclinic now (full-size cells); x and y shrank below 5 list cutoffs
    assert list(mgr.stats()["n_cells"]) == [5, 5, 5]    # ... and by step 70 the box has bounced back
    assert np.abs(np.diagonal(hh, axis1=1, axis2=2) / np.diag(hh[0]) - 1.0).max() > 0.05


def test_cuda_graph_batches_equal_classic_launches():
    """pisb_step_nve replays CUDA graphs of 8 / 4 / 2 steps with the rebuild chain inside a device-side conditional node;
    the classic per-kernel launch sequence (option cuda_graphs = 0) must give bit-identical traces and states, across
    rebuilds, odd step counts (the f/g force buffers trade places) and repeated calls."""
    atoms_a = fcc_argon(12, temperature=60.0, seed=5)      # hot: several rebuilds in 150 steps
    atoms_b = fcc_argon(12, temperature=60.0, seed=5)
    res = []
    for atoms, use in ((atoms_a, 1), (atoms_b, 0)):
        mgr = make_manager(skin=SKIN)
        mgr.set_option("cuda_graphs", use)
        mgr.attach(atoms)
        mgr.compute()
        th = [mgr.step_nve(0.25, k) for k in (64, 7, 1, 33, 46)]   # multiples of 8, odd remainders (f/g parity flips), a single step
        st = mgr.stats()
        mgr.download(atoms)
        res.append((np.concatenate(th), st))
    (tg, sg), (tc, sc) = res
    for key in ("pe", "ke", "virial_ref", "virial_pair"):
        assert np.array_equal(tg[key], tc[key]), key
    assert np.array_equal(atoms_a.positions, atoms_b.positions) and np.array_equal(atoms_a.velocities, atoms_b.velocities)
    assert np.array_equal(atoms_a.forces, atoms_b.forces)
    assert sg["n_builds"] == sc["n_builds"] and sg["n_builds"] >= 4
    assert sg["n_launches"] < sc["n_launches"]          # skipped rebuild chains are not launched at all


def test_cuda_graph_small_system_matches_oracle():
    """The reference's own example scale (4000 atoms, example/argon4000.txt geometry) through graph replays vs the oracle."""
    atoms = fcc_argon(10, temperature=30.0, seed=2)
    orc = make_oracle(atoms, {(1, 1): argon_pair(8.5)})
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref = orc.run_nve(x, v, np.zeros_like(x), atoms.type_ids, 0.25, 200)
    mgr = make_manager(skin=SKIN, rc=8.5)
    mgr.attach(atoms)
    mgr.compute()
    th = mgr.step_nve(0.25, 200)
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL


def _run_batches(atoms, fuse, graphs, batches, table=None, nvt_first=False, variant=6):
    mgr = make_manager(skin=SKIN, variant=variant, table=table)      # force_variant 3 / 5: the fused step is selectable at any size
    mgr.set_option("fuse_vv", fuse)
    mgr.set_option("cuda_graphs", graphs)
    mgr.attach(atoms)
    mgr.compute()
    if nvt_first:   # an NVT batch flips the force buffers but not the position buffers: the NVE graphs must cope
        chain = mgr.nhc_new(60.0, 60.0, 50.0)
        mgr.step_nvt_nhc(0.25, 5, chain, 0, 5)
    th = np.concatenate([mgr.step_nve(0.25, k) for k in batches])
    st = mgr.stats()
    mgr.download(atoms)
    mgr.close()
    return th, st


@pytest.mark.parametrize("variant", [6, 11])
@pytest.mark.parametrize("graphs", [1, 0])
def test_fused_force_integrator_step_equals_separate_kernels(graphs, variant):
    """k_force_vv (force + kick + drift in one launch, positions double-buffered) against k_force_v3 + k_vv: the same
    arithmetic per atom, so positions / velocities / forces and the force kernel's reductions (PE, pair virial) are
    bit-identical across rebuilds, odd batch sizes and single steps; KE and tr(X F^T) are reduced over 512- instead of
    256-thread blocks and agree to rounding."""
    batches = (64, 7, 1, 33, 46, 2, 3)
    a1 = fcc_argon(12, temperature=60.0, seed=5)
    a2 = fcc_argon(12, temperature=60.0, seed=5)
    t1, s1 = _run_batches(a1, 1, graphs, batches, variant=variant)
    t0, s0 = _run_batches(a2, 0, graphs, batches, variant=variant)
    assert np.array_equal(a1.positions, a2.positions)
    assert np.array_equal(a1.velocities, a2.velocities)
    assert np.array_equal(a1.forces, a2.forces)
    for key in ("pe", "virial_pair"):
        assert np.array_equal(t1[key], t0[key]), key
    for key in ("ke", "virial_ref"):
        assert np.max(np.abs(t1[key] - t0[key]) / np.maximum(np.abs(t0[key]), 1e-3)) <= 1e-12, key
    assert s1["n_builds"] == s0["n_builds"] and s1["n_builds"] >= 4
    assert s1["n_launches"] < s0["n_launches"]     # one kernel per step instead of two


def test_fused_step_after_an_nvt_batch_and_with_two_types():
    """Buffer-parity corner: an NVT batch (unfused) flips f/g only, the fused NVE batches that follow flip f/g and the
    position buffers together.  Two atom types exercise the table path of k_force_vv."""
    table = {(1, 1): LennardJones(0.238, 3.405, 8.5), (1, 2): LennardJones(0.15, 3.0, 7.5), (2, 2): LennardJones(0.07, 2.8, 7.0)}
    res = []
    for fuse in (1, 0):
        atoms = fcc_argon(10, temperature=60.0, seed=9)
        atoms.type_ids[::3] = 2
        atoms.masses = [39.948, 20.18]
        th, st = _run_batches(atoms, fuse, 1, (8, 5, 16, 1, 10), table=table, nvt_first=True)
        res.append((atoms, th, st))
    (a1, t1, s1), (a0, t0, s0) = res
    assert np.array_equal(a1.positions, a0.positions) and np.array_equal(a1.velocities, a0.velocities)
    assert np.array_equal(a1.forces, a0.forces)
    assert np.array_equal(t1["pe"], t0["pe"])
    assert np.max(np.abs(t1["ke"] - t0["ke"]) / np.abs(t0["ke"])) <= 1e-12
    assert s1["n_builds"] == s0["n_builds"]


def test_fused_step_trace_matches_oracle():
    """The fused step against the oracle directly (the 256k- and 4M-atom tests reach it through the defaults; this one
    forces it at 4000 atoms for 300 steps with rebuilds)."""
    atoms = fcc_argon(10, temperature=40.0, seed=21)
    orc = make_oracle(atoms, {(1, 1): argon_pair(8.5)})
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    f = np.zeros_like(x)
    ref = orc.run_nve(x, v, f, atoms.type_ids, 0.25, 300)
    mgr = make_manager(skin=SKIN, rc=8.5, variant=6)
    mgr.attach(atoms)
    mgr.compute()
    th = mgr.step_nve(0.25, 300)
    assert mgr.stats()["n_builds"] >= 3
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL
    assert_virial_trace(atoms, th, ref, x, f)
    mgr.download(atoms)
    assert np.abs(atoms.positions - x).max() < 1e-8


@pytest.mark.parametrize("variant", [0, 6, 11])
def test_pipelined_host_step_equals_whole_array_step(variant):
    """pisb_verlet_step_nve_host cuts the trait call's 3 x N host arrays into chunks and pipelines upload, drift and
    download (k_host_load_drift / k_store_range); option host_pipeline = 0 keeps the whole-array sequence.  Same
    arithmetic per atom, so 40 hot steps (with rebuilds, ragged last chunk) are bit-identical -- with the separate
    kernels (variant 0 at this size) and with the fused force + kick kernel (variant 6)."""
    res = []
    for pipe in (1, 0):
        atoms = fcc_argon(9, temperature=80.0, seed=17)            # 2916 atoms
        mgr = make_manager(skin=SKIN, variant=variant)
        mgr.set_option("host_pipeline", pipe)
        mgr.set_option("host_chunk_atoms", 700)                    # 5 chunks, the last one ragged
        mgr.compute_potential(atoms)
        pes = [mgr.verlet_step_nve(atoms, 0.25) for _ in range(40)]
        res.append((atoms, np.array(pes), mgr.stats()))
        mgr.close()
    (a1, p1, s1), (a0, p0, s0) = res
    assert np.array_equal(a1.positions, a0.positions)
    assert np.array_equal(a1.velocities, a0.velocities)
    assert np.array_equal(a1.forces, a0.forces)
    assert np.array_equal(p1, p0)
    assert s1["n_builds"] == s0["n_builds"] and s1["n_builds"] >= 2


def test_pipelined_host_step_follows_host_side_edits():
    """The host arrays are authoritative at every call: velocities rescaled and an atom moved by the caller between two
    trait calls must be honoured (the list is rebuilt if the edit broke the skin criterion)."""
    atoms = fcc_argon(8, temperature=30.0, seed=4)
    table = {(1, 1): argon_pair()}
    mgr = make_manager(skin=SKIN)
    mgr.compute_potential(atoms)
    for _ in range(3):
        mgr.verlet_step_nve(atoms, 0.25)
    atoms.velocities *= 0.5
    atoms.positions[5] += np.array([1.2, -0.5, 0.4])               # beyond skin/2 = 0.51
    orc = make_oracle(atoms, table)
    x, v, f = atoms.positions.copy(), atoms.velocities.copy(), atoms.forces.copy()
    pe_ref = orc.verlet_step_nve(x, v, f, atoms.type_ids, 0.25)
    pe = mgr.verlet_step_nve(atoms, 0.25)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert np.array_equal(atoms.positions, x)
    assert force_rel_err(atoms.forces, f).max() <= FORCE_TOL
    assert np.abs(atoms.velocities - v).max() <= 1e-12 * max(1.0, np.abs(v).max())


def test_asynchronous_download_is_a_snapshot():
    """pisb_download_begin snapshots the state in stream order; steps enqueued afterwards must not leak into it."""
    from pis_b200 import capi

    atoms = fcc_argon(8, temperature=50.0, seed=8, pinned=True)
    mgr = make_manager(skin=SKIN)
    mgr.attach(atoms)
    mgr.compute()
    mgr.step_nve(0.25, 20)
    ref = fcc_argon(8, temperature=50.0, seed=8)
    mgr.download(ref)
    mgr.download_begin(atoms, positions=True, velocities=True, forces=True)
    with pytest.raises(capi.PisbError):
        mgr.download_begin(atoms)                     # one download in flight per handle
    mgr.step_nve(0.25, 30)                            # the next batch runs while the copy is in flight
    mgr.download_end()
    assert np.array_equal(atoms.positions, ref.positions)
    assert np.array_equal(atoms.velocities, ref.velocities)
    assert np.array_equal(atoms.forces, ref.forces)
    mgr.download_end()                                # idempotent
    # page-locking caller-owned arrays (what the Rust shim does with nalgebra's storage)
    buf = np.zeros((atoms.n_atoms, 3))
    lib = capi.load()
    assert lib.pisb_host_register(capi._ptr(buf), buf.nbytes) == capi.PISB_OK
    plain = fcc_argon(8, temperature=50.0, seed=8)
    plain.positions = buf
    mgr.download_begin(plain, positions=True)
    mgr.download_end()
    assert lib.pisb_host_unregister(capi._ptr(buf)) == capi.PISB_OK
    mgr.download(ref)
    assert np.array_equal(buf, ref.positions)


@pytest.mark.parametrize("two_types", [False, True])
def test_device_velocity_initialisation_matches_the_host_generator(two_types):
    """pisb_start_velocities = Atoms::start_velocities (velocities.rs:10-59) on the device: Gaussian components with
    sigma_i = sqrt(kB T / m_i) from the repository's id-keyed generator, remove_drift, rescale_to_temperature.  Against the
    host implementation of the same three steps (pis_b200.lattice.create_velocities, what the oracle runs are fed with):
    equal to rounding (device log / cos differ from libm by an ulp), momentum zero, temperature exact; the state the
    handle holds otherwise (positions, list, forces) is untouched, and NVE steps from the device-made velocities follow
    the oracle's from the host-made ones."""
    from pis_b200.atoms import KB_KJPERMOLEKELVIN
    from pis_b200.lattice import create_velocities

    atoms = fcc_argon(9, temperature=0.0, seed=3, jitter=0.05)
    n = atoms.n_atoms
    table = {(1, 1): argon_pair()}
    if two_types:
        types = np.where(np.arange(n) % 3 == 0, 2, 1).astype(np.int32)
        atoms = Atoms(types, [ARGON["mass"], 83.798], atoms.positions, atoms.sim_box, velocities=atoms.velocities)
        table = {(1, 1): argon_pair(), (1, 2): LennardJones(0.3, 3.5, 8.0, True), (2, 2): LennardJones(0.4, 3.6, 8.2, True)}
    m_per_atom = np.asarray(atoms.masses)[atoms.type_ids - 1]
    mgr = make_manager(skin=SKIN, table=table)
    mgr.attach(atoms)
    pe0 = mgr.compute()
    mgr.start_velocities(37.5, 2024)
    mgr.download(atoms, positions=False, forces=False)
    v_host = create_velocities(n, m_per_atom, 37.5, 2024)
    assert np.abs(atoms.velocities - v_host).max() <= 1e-13 * np.abs(v_host).max()
    p = (atoms.velocities * m_per_atom[:, None]).sum(axis=0)
    assert np.abs(p).max() <= 1e-9 * np.abs(atoms.velocities * m_per_atom[:, None]).sum()
    ke = mgr.thermo_now()["ke"]
    assert abs(2.0 * ke / (3.0 * n * KB_KJPERMOLEKELVIN) - 37.5) <= 1e-12 * 37.5
    # the run that follows: device-made velocities on the GPU, host-made ones in the oracle
    orc = make_oracle(atoms, table)
    x, v, f = atoms.positions.copy(), v_host.copy(), np.zeros_like(v_host)
    ref = orc.run_nve(x, v, f, atoms.type_ids, 0.25, 40)
    assert abs(pe0 - ref[0, 0]) <= ENERGY_TOL * abs(ref[0, 0])
    th = mgr.step_nve(0.25, 40)
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL
    with pytest.raises(capi.PisbError):
        mgr.start_velocities(-1.0, 1)     # sqrt of a negative variance: InvalidDistribution (velocities.rs:25-26)


@pytest.mark.parametrize("variant", [4, 6, 7, 0])
def test_multi_type_lists_exact_with_every_build(variant):
    """Per-pair cutoffs in the list build: three types, six different rc + skin radii, one pair missing
    from the table (never listed, lennard_jones.rs:216-222 skips it), on a box large enough for interior warps (the packed
    FP32 v3 build without the image search) and boundary warps alike.  Rows exact against the oracle for the scalar
    pre-filter build (variant 4) and the packed-FP32 v3 build with its shared-memory band table (6, 7, and the default)."""
    atoms = fcc_argon(12, temperature=30.0, seed=5, jitter=0.25)
    n = atoms.n_atoms
    types = (1 + (np.arange(n) * 7 + (np.arange(n) // 5)) % 3).astype(np.int32)
    atoms = Atoms(types, [39.948, 20.18, 83.798], atoms.positions, atoms.sim_box, velocities=atoms.velocities)
    table = {(1, 1): LennardJones(0.238, 3.405, 8.5), (1, 2): LennardJones(0.15, 3.0, 7.5), (2, 2): LennardJones(0.07, 2.8, 7.0),
             (1, 3): LennardJones(0.3, 3.5, 9.0), (3, 3): LennardJones(0.4, 3.6, 6.2)}          # (2, 3) missing
    orc = make_oracle(atoms, table)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids)
    mgr = make_manager(skin=SKIN, table=table, variant=variant)
    mgr.attach(atoms)
    pe = mgr.compute()
    rows = mgr.neighbours(n)
    ref_rows = csr_rows_sorted(start, nbr)
    bad = [i for i in range(n) if not np.array_equal(rows[i], ref_rows[i])]
    assert not bad, f"{len(bad)} rows differ, first atom {bad[0]}"
    mgr.download(atoms, positions=False, velocities=False)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
    # and a short hot run with rebuilds on the same table.  dt = 0.1: the jittered three-type lattice starts far up its
    # repulsive walls (the light type 2 atoms reach ~60 K within ten steps) and dt = 0.25 is beyond the integrator's
    # stability limit there -- the oracle's own trace overflows to NaN after ~25 steps, which pins nothing.
    x, v, f = atoms.positions.copy(), atoms.velocities.copy(), np.zeros_like(atoms.positions)
    ref = orc.run_nve(x, v, f, atoms.type_ids, 0.1, 40)
    assert np.isfinite(ref).all()
    builds_before = mgr.stats()["n_builds"]
    th = mgr.step_nve(0.1, 40)
    assert mgr.stats()["n_builds"] >= builds_before + 3
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= ENERGY_TOL


def test_list_capacity_is_raised_ahead_of_an_overflow_in_a_compressing_box():
    """A build inside a batch cannot reallocate, so the capacity must stay AHEAD of the rows (grow_list_if_close: raised as
    soon as the longest row comes within 12 % of it, nothing truncated yet).  Scenario: an expanded lattice (a = 6.0) under
    NPT collapses by 45 % in volume within 130 steps -- the rows grow from ~67 to ~120 entries, past the 112 slots the
    initial density suggested.  The run must go through without PISB_ERR_CAPACITY, raise the capacity on the way, and
    still follow the oracle's NPT trace."""
    atoms = fcc_argon(7, temperature=5.0, seed=12345, a=6.0)
    table = {(1, 1): argon_pair(8.5)}
    orc = make_oracle(atoms, table)
    chain_ref = orc.nhc_new(5.0, 50.0, 100.0)
    baro_ref = orc.mtk_new(0.02, 100.0, atoms.n_atoms, 5.0)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    steps = 130
    ref, htr = orc.run_npt(x, v, np.zeros_like(x), atoms.type_ids, 0.25, steps, baro_ref, chain_ref)
    mgr = make_manager(skin=SKIN, rc=8.5)
    mgr.attach(atoms)
    mgr.compute()
    cap0 = mgr.stats()["list_capacity"]
    chain = mgr.nhc_new(5.0, 50.0, 100.0)
    baro = mgr.mtk_new(0.02, 100.0, atoms.n_atoms, 5.0)
    th, en, hh = mgr.step_npt_mtk(0.25, steps, baro, chain, 0, steps, atoms)
    st = mgr.stats()
    assert st["capacity_growths"] >= 1 and st["list_capacity"] > cap0 and st["max_neighbours"] > cap0
    vol = np.abs(np.linalg.det(hh))
    assert vol[-1] < 0.65 * vol[0]
    assert np.max(np.abs(th["pe"] - ref[1:, 0]) / np.abs(ref[1:, 0])) <= 1e-8
    assert np.max(np.abs(th["ke"] - ref[1:, 1]) / np.abs(ref[1:, 1])) <= 1e-8
    h_ref = htr[1:].reshape(-1, 3, 3).transpose(0, 2, 1)
    assert np.abs(hh - h_ref).max() <= 1e-8 * np.abs(h_ref).max()
