"""GPU parity on the configurations the numbers are quoted on (BASELINE.json configs), at full size
against the CPU oracle (OpenMP variant where the serial loop would take minutes; same per-pair arithmetic):

  C1  the reference's own example/input.pis + example/argon4000.txt through the host CLI (as shipped = NPT, and NVE)
  C2  256 000 atoms x 1000 NVE steps at T0 = 5 K                      (PE / KE / T traces 1e-9)
  C3  4 000 000 atoms, thermally displaced, one shot                  (forces 1e-10, neighbour sets exact)
  C5  rc in {3, 4, 5} sigma x rho* in {0.6, 1.0}                      (sets exact, forces 1e-10; list capacity regrows)
  +   the 108-atom fixture against EXACT arithmetic (tests/golden/exact_small.json)

Bars: neighbour lists bit-exact as sorted sets; forces 1e-10 relative; energy / temperature traces 1e-9 relative."""
import json
import os
import subprocess

import numpy as np
import pytest

from pis_b200 import Atoms, LennardJones, SimulationBox
from pis_b200.lattice import ARGON, create_velocities, fcc_argon
from tests.example_inputs import rust_display, write_example
from tests.helpers import SIGMA, SKIN, argon_pair, force_rel_err, make_manager, make_oracle

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-10
ENERGY_TOL = 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "pis_b200", "pis_b200_cli")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rows_equal_csr(counts, rows, start, nbr, chunk=1 << 19):
    """(counts, padded rows) from the GPU == CSR (start, nbr) from the oracle as SORTED sets, vectorised in chunks.
    Returns the number of atoms whose rows differ."""
    n = len(counts)
    ref_counts = np.diff(start)
    bad = int((ref_counts != counts).sum())
    if bad:
        return bad
    width = rows.shape[1]
    big = np.iinfo(np.int32).max
    col = np.arange(width, dtype=np.int64)[None, :]
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        c = counts[a:b].astype(np.int64)[:, None]
        mask = col < c
        g = np.where(mask, rows[a:b], big)
        g.sort(axis=1)
        idx = np.minimum(start[a:b, None] + col, len(nbr) - 1)
        r = np.where(mask, nbr[idx], big)
        r.sort(axis=1)
        bad += int((g != r).any(axis=1).sum())
    return bad


# ---------------------------------------------------------------------------------------------------------------------
# C3: 4M atoms, one shot
# ---------------------------------------------------------------------------------------------------------------------
def test_config3_4m_atoms_forces_and_lists_vs_oracle():
    """BASELINE configs[2] at FULL size: FCC argon 100^3 x 4 = 4 000 000 atoms, rc = 2.5 sigma, skin = 0.3 sigma, positions
    thermally displaced (Gaussian, 0.1 A) so forces do not cancel.  compute_potential (lennard_jones.rs:186-244) vs the
    oracle's all-core loop: PE 1e-9, per-atom forces 1e-10; the Verlet list (lennard_jones.rs:345-415 with rcut + skin) vs
    the oracle's list for ALL 4M atoms as sorted sets, exact."""
    atoms = fcc_argon(100, temperature=43.0, seed=12345, jitter=0.1)
    n = atoms.n_atoms
    assert n == 4_000_000
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids, mode="omp")
    mgr = make_manager(skin=SKIN)
    mgr.attach(atoms)
    pe = mgr.compute()
    mgr.download(atoms, positions=False, velocities=False)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    err = force_rel_err(atoms.forces, f_ref)
    assert err.max() <= FORCE_TOL, err.max()
    del f_ref, err
    counts, rows = mgr.neighbours_padded(n)
    mgr.close()
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN, mode="omp", cap_per_atom=int(counts.max()) + 2)
    assert len(nbr) == int(counts.astype(np.int64).sum())
    assert rows_equal_csr(counts, rows, start, nbr) == 0


# ---------------------------------------------------------------------------------------------------------------------
# C2: 256k atoms x 1000 steps
# ---------------------------------------------------------------------------------------------------------------------
def test_config2_256k_atoms_1000_nve_steps_vs_oracle():
    """BASELINE configs[1] verbatim: FCC argon 40^3 x 4 = 256 000 atoms, rc = 2.5 sigma, skin = 0.3 sigma, dt = 0.25,
    T0 = 5 K, 1000 NVE steps.  PE / KE / temperature of EVERY step vs the oracle's run (all-core loop) within 1e-9
    relative; positions after 1000 steps within 1e-9 A."""
    atoms = fcc_argon(40, temperature=5.0, seed=12345)
    table = {(1, 1): argon_pair()}
    orc = make_oracle(atoms, table)
    steps = 1000
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    f = np.zeros_like(x)
    ref = orc.run_nve(x, v, f, atoms.type_ids, 0.25, steps, mode="omp")
    mgr = make_manager(skin=SKIN)
    mgr.attach(atoms)
    pe0 = mgr.compute()
    assert abs(pe0 - ref[0, 0]) <= ENERGY_TOL * abs(ref[0, 0])
    th = mgr.step_nve(0.25, steps)
    assert mgr.stats()["n_builds"] >= 1
    pe_ref, ke_ref, t_ref = ref[1:, 0], ref[1:, 1], ref[1:, 3]
    assert np.max(np.abs(th["pe"] - pe_ref) / np.abs(pe_ref)) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - ke_ref) / np.abs(ke_ref)) <= ENERGY_TOL
    t_gpu = np.array([atoms.temerature(k) for k in th["ke"]])
    assert np.max(np.abs(t_gpu - t_ref) / np.abs(t_ref)) <= ENERGY_TOL
    mgr.download(atoms)
    assert np.abs(atoms.positions - x).max() <= 1e-9
    # forces of the two runs differ by the trajectories' accumulated rounding (|dx| ~ 1e-10 A x the LJ stiffness): compare them
    # on ONE state instead -- the oracle's final positions, evaluated by both
    assert force_rel_err(atoms.forces, f).max() <= 1e-8
    atoms.positions[...] = x
    mgr.attach(atoms)
    mgr.compute()
    mgr.download(atoms, positions=False, velocities=False)
    assert force_rel_err(atoms.forces, f).max() <= FORCE_TOL


# ---------------------------------------------------------------------------------------------------------------------
# C5: cutoff / density sweep
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rc_sigma", [3.0, 4.0, 5.0])
@pytest.mark.parametrize("rho", [0.6, 1.0])
def test_config5_cutoff_density_sweep_vs_oracle(rc_sigma, rho):
    """BASELINE configs[4] points (rc = 3 / 4 / 5 sigma, rho* = 0.6 / 1.0, a = sigma (4 / rho*)^(1/3)) on 32 000 atoms:
    K = 100 ... 620 neighbours per atom, so the list capacity estimate is exceeded and regrown.  Sets exact, forces 1e-10,
    then 20 NVE steps (PE / KE 1e-9)."""
    a = SIGMA * (4.0 / rho) ** (1.0 / 3.0)
    rc = rc_sigma * SIGMA
    atoms = fcc_argon(20, temperature=43.0, seed=99, jitter=0.12, a=a)
    table = {(1, 1): argon_pair(rc)}
    orc = make_oracle(atoms, table)
    pe_ref, f_ref = orc.compute_potential(atoms.positions, atoms.type_ids, mode="omp")
    mgr = make_manager(skin=SKIN, rc=rc)
    mgr.attach(atoms)
    counts, rows = mgr.neighbours_padded(atoms.n_atoms)
    start, nbr = orc.build_neighbour_list(atoms.positions, atoms.type_ids, extra=SKIN, mode="omp", cap_per_atom=int(counts.max()) + 2)
    assert rows_equal_csr(counts, rows, start, nbr) == 0
    k_mean = counts.mean()
    assert 0.8 < k_mean / (4.18879 * (rc + SKIN) ** 3 * 4.0 / a ** 3) < 1.2      # K ~ 4/3 pi (rc + skin)^3 rho
    pe = mgr.compute()
    mgr.download(atoms, positions=False, velocities=False)
    assert abs(pe - pe_ref) <= ENERGY_TOL * abs(pe_ref)
    assert force_rel_err(atoms.forces, f_ref).max() <= FORCE_TOL
    st = mgr.stats()
    assert st["list_capacity"] >= st["max_neighbours"] >= counts.max()
    x, v, f = atoms.positions.copy(), atoms.velocities.copy(), f_ref.copy()
    pes, kes = [], []
    for _ in range(20):
        pes.append(orc.verlet_step_nve(x, v, f, atoms.type_ids, 0.25, mode="omp"))
        kes.append(orc.kinetic_energy(v, atoms.type_ids))
    th = mgr.step_nve(0.25, 20)
    assert np.max(np.abs(th["pe"] - np.array(pes)) / np.abs(pes)) <= ENERGY_TOL
    assert np.max(np.abs(th["ke"] - np.array(kes)) / np.abs(kes)) <= ENERGY_TOL


# ---------------------------------------------------------------------------------------------------------------------
# exact arithmetic
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1, 3, 8])
def test_gpu_matches_exact_arithmetic(variant):
    """The 108-atom fixture against a 60-digit evaluation of the reference's formulas (tests/golden/make_exact.py):
    PE, forces, the neighbour sets, and one verlet_step_nve with its observables -- the GPU path held to the formulas of
    lennard_jones.rs:33-55 / potential.rs:15-33 directly, not to any other implementation."""
    with open(os.path.join(GOLDEN, "oracle_small.json")) as fh:
        g = json.load(fh)
    with open(os.path.join(GOLDEN, "exact_small.json")) as fh:
        e = json.load(fh)
    L = g["L"]
    box = SimulationBox.from_lammps_data(0, L, 0, L, 0, L)
    atoms = Atoms(np.array(g["types"], dtype=np.int32), [g["mass"]], np.array(g["positions"]), box,
                  velocities=np.array(g["velocities"]))
    table = {(1, 1): LennardJones(g["eps"], g["sigma"], g["rc"], True)}
    mgr = make_manager(skin=g["skin"], table=table, variant=variant)
    mgr.attach(atoms)
    rows = mgr.neighbours(atoms.n_atoms)
    assert [r.tolist() for r in rows] == e["neighbours_skin"]
    pe0 = mgr.compute()
    mgr.download(atoms, positions=False, velocities=False)
    f0 = np.array(e["forces0"])
    fscale = np.sqrt((f0 ** 2).sum(axis=1).mean())
    assert abs(pe0 - e["pe0"]) <= 1e-12 * abs(e["pe0"])
    assert np.abs(atoms.forces - f0).max() <= 1e-12 * fscale
    th = mgr.step_nve(g["dt"], 1)
    mgr.download(atoms)
    assert np.abs(atoms.positions - np.array(e["positions1"])).max() <= 1e-13 * L
    assert np.abs(atoms.velocities - np.array(e["velocities1"])).max() <= 1e-12 * np.abs(np.array(g["velocities"])).max()
    assert np.abs(atoms.forces - np.array(e["forces1"])).max() <= 1e-12 * fscale
    assert abs(th["pe"][0] - e["pe1"]) <= 1e-12 * abs(e["pe1"])
    assert abs(th["ke"][0] - e["ke1"]) <= 1e-12 * abs(e["ke1"])
    # sum r.F cancels heavily: its error is bounded by rounding of the terms, i.e. relative to sum |r.F|
    assert abs(th["virial_ref"][0] - e["virial_trace1"]) <= 1e-12 * np.abs(atoms.positions * atoms.forces).sum()


# ---------------------------------------------------------------------------------------------------------------------
# C1: the reference's own example files through the host CLI
# ---------------------------------------------------------------------------------------------------------------------
def _cli_thermo(tmp_path, nve, steps):
    write_example(str(tmp_path), nve=nve, steps=steps)
    r = subprocess.run([CLI, "-i", "input.pis", "--skin", repr(SKIN)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    out = r.stdout.strip().splitlines()
    assert len(out) == steps + 1
    return out


def _example_state():
    atoms = fcc_argon(10, temperature=0.0)
    m = np.full(atoms.n_atoms, ARGON["mass"])
    # `velocity all create 5.0 12345` after read_data (example/input.pis:22): this repo's id-keyed generator (the
    # reference's rand::SmallRng stream is third-party and unpinned) + remove_drift + rescale (velocities.rs:35-59)
    atoms.velocities[...] = create_velocities(atoms.n_atoms, m, 5.0, 12345)
    return atoms


def test_config1_reference_example_nve_through_the_cli(tmp_path):
    """BASELINE configs[0]: example/argon4000.txt (the reference's bytes, regenerated and sha-checked) driven by
    example/input.pis with its `fix ... npt` line removed (NVE), 200 steps through pis_b200_cli: every thermo
    line (printed {:.3}) and every dump frame vs the oracle's Simulation::run."""
    from oracle.pis_oracle import Oracle

    steps = 200
    out = _cli_thermo(tmp_path, True, steps)
    atoms = _example_state()
    n, L = atoms.n_atoms, 54.1
    o = Oracle.cubic(L)
    o.insert(1, 1, 0.238, 3.405, 8.5)               # the data file's PairCoeffs line
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref = o.run_nve(x, v, np.zeros_like(x), atoms.type_ids, 0.25, steps)
    assert abs(float(out[0].split()[1]) - (-6956.99645673589)) < 1e-8      # step-0 PE of the lattice, full precision
    for s in range(1, steps + 1):
        got = [float(t) for t in out[s].split()[1:]]
        want = [ref[s, 0], ref[s, 1], ref[s, 0] + ref[s, 1], ref[s, 3], ref[s, 4]]
        for g_, e_ in zip(got, want):
            assert abs(g_ - e_) <= 0.51e-3 + 1e-9 * abs(e_), (s, got, want)
    dump = (tmp_path / "dump.lammpstrj").read_text().splitlines()
    frame = 9 + n
    assert len(dump) == frame * (steps // 10 + 1)           # `dump dp1 all atom 10`: frame 0 + every 10th step
    assert dump[5] == "0 " + rust_display(L) and dump[9] == "1 1 0 0 0" and dump[10] == "2 1 0 2.705 2.705"
    rows = np.array([[float(t) for t in ln.split()] for ln in dump[-n:]])
    assert np.abs(rows[:, 2:] - x).max() < 1e-9


def test_config1_reference_example_as_shipped_npt_through_the_cli(tmp_path):
    """example/input.pis AS SHIPPED (`fix mynpt all npt temp 5.0 50.0 100 iso 0.01 0.01 1000`, potential.rs:112-135) on
    example/argon4000.txt, `run` shortened to 200 steps: thermo lines and the final dump frame vs the oracle's NPT run."""
    from oracle.pis_oracle import Oracle

    steps = 200
    out = _cli_thermo(tmp_path, False, steps)
    atoms = _example_state()
    n, L = atoms.n_atoms, 54.1
    o = Oracle.cubic(L)
    o.insert(1, 1, 0.238, 3.405, 8.5)
    chain = o.nhc_new(5.0, 50.0, 100.0)
    baro = o.mtk_new(0.01, 1000.0, n, 5.0)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref, htr = o.run_npt(x, v, np.zeros_like(x), atoms.type_ids, 0.25, steps, baro, chain)
    for s in range(1, steps + 1):
        got = [float(t) for t in out[s].split()[1:]]
        for g_, e_ in zip(got, ref[s]):
            assert abs(g_ - e_) <= 0.51e-3 + 1e-9 * abs(e_), (s, got, list(ref[s]))
    dump = (tmp_path / "dump.lammpstrj").read_text().splitlines()
    frame = 9 + n
    assert len(dump) == frame * (steps // 10 + 1)
    hs = htr[steps].reshape(3, 3).T
    last = dump[-frame:]
    for d in range(3):
        assert abs(float(last[5 + d].split()[1]) - hs[d, d]) <= 1e-9 * hs[d, d]
    rows = np.array([[float(t) for t in ln.split()] for ln in last[9:]])
    assert np.abs(rows[:, 2:] - x).max() < 1e-8
    assert abs(hs[0, 0] - L) > 1e-6
