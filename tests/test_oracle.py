"""CPU tests of the oracle (oracle/pis_oracle.c).  The reference's own tests pin no physics and the
Rust reference cannot be built here, so the oracle is pinned against analytic known answers, an
independent numpy restatement, its own O(N^2) driver and the committed golden vectors."""
import json
import os

import numpy as np
import pytest

from oracle.pis_oracle import Oracle
from pis_b200.lattice import fcc_argon

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
EPS, SIG = 0.238, 3.405


def numpy_lj(pos, L, eps, sig, rc):
    """Independent restatement: O(N^2), minimum image as d - L*round(d/L) (NOT the reference's h*(hinv*d) form)."""
    n = len(pos)
    d = pos[None, :, :] - pos[:, None, :]
    d -= L * np.round(d / L)
    r2 = (d * d).sum(-1)
    np.fill_diagonal(r2, np.inf)
    mask = np.sqrt(r2) <= rc
    inv = np.where(mask, 1.0 / r2, 0.0)
    s6 = (sig * sig * inv) ** 3
    ucut = 4 * eps * ((sig / rc) ** 12 - (sig / rc) ** 6)
    u = np.where(mask, 4 * eps * (s6 * s6 - s6) - ucut, 0.0)
    fs = np.where(mask, 24 * eps * (2 * s6 * s6 - s6) * inv, 0.0)
    f = -(fs[:, :, None] * d).sum(axis=1)
    return 0.5 * u.sum(), f, mask


def make(L, rc=8.5):
    o = Oracle.cubic(L)
    o.insert(1, 1, EPS, SIG, rc)
    return o


def test_two_atom_minimum():
    rc = 8.5
    o = make(40.0, rc)
    r = 2.0 ** (1.0 / 6.0) * SIG
    u, f = o.lj_pair(EPS, SIG, rc, [r, 0.0, 0.0])
    ucut = 4 * EPS * ((SIG / rc) ** 12 - (SIG / rc) ** 6)
    assert abs(u - (-EPS - ucut)) < 1e-14
    assert np.abs(f).max() < 1e-13
    u2, _ = o.lj_pair(EPS, SIG, rc, [r, 0.0, 0.0], shift=False)
    assert abs(u2 + EPS) < 1e-14


def test_argon4000_lattice_sum():
    """example/argon4000.txt geometry (10^3 FCC cells, a = 5.41, rc = 8.5): SURVEY 8c known answer."""
    atoms = fcc_argon(10, temperature=0.0)
    o = make(54.1)
    assert o.divide_into_cells(o.max_rcut()) == (6, 6, 6)
    pe, f = o.compute_potential(atoms.positions, atoms.type_ids)
    assert abs(pe - (-6956.99645673589)) < 1e-8
    assert np.abs(f).max() < 1e-12
    start, nbr = o.build_neighbour_list(atoms.positions, atoms.type_ids)
    assert len(nbr) == 216000 and np.all(np.diff(start) == 54)
    start, nbr = o.build_neighbour_list(atoms.positions, atoms.type_ids, extra=0.3 * SIG)
    assert np.all(np.diff(start) == 86)


def test_cell_driver_equals_n2_and_numpy():
    atoms = fcc_argon(6, temperature=0.0, jitter=0.25, seed=4)
    L = atoms.sim_box.h[0, 0]
    o = make(L)
    pe, f = o.compute_potential(atoms.positions, atoms.type_ids)
    pe2, f2 = o.compute_potential(atoms.positions, atoms.type_ids, mode="n2")
    pe3, f3 = o.compute_potential(atoms.positions, atoms.type_ids, mode="omp", threads=3)
    assert abs(pe - pe2) < 1e-11 * abs(pe) and np.abs(f - f2).max() < 1e-12
    assert abs(pe - pe3) < 1e-11 * abs(pe) and np.abs(f - f3).max() < 1e-12
    pe_np, f_np, _ = numpy_lj(atoms.positions, L, EPS, SIG, 8.5)
    assert abs(pe - pe_np) < 1e-10 * abs(pe)
    assert np.abs(f - f_np).max() < 1e-10 * np.abs(f_np).max()
    assert np.abs(f.sum(axis=0)).max() < 1e-11  # Newton's third law


def test_full_list_matches_bruteforce_and_list_consumer():
    atoms = fcc_argon(6, temperature=0.0, jitter=0.3, seed=9)
    L = atoms.sim_box.h[0, 0]
    o = make(L)
    for extra in (0.0, 1.0215):
        start, nbr = o.build_neighbour_list(atoms.positions, atoms.type_ids, extra=extra)
        _, _, mask = numpy_lj(atoms.positions, L, EPS, SIG, 8.5 + extra)
        for i in range(0, atoms.n_atoms, 37):
            assert np.array_equal(np.sort(nbr[start[i]:start[i + 1]]), np.nonzero(mask[i])[0])
        pe_l, f_l = o.compute_potential_list(atoms.positions, atoms.type_ids, start, nbr)
        pe, f = o.compute_potential(atoms.positions, atoms.type_ids)
        assert abs(pe - pe_l) < 1e-11 * abs(pe) and np.abs(f - f_l).max() < 1e-12


def test_forward_offsets_are_a_half_shell():
    import ctypes as C

    o = make(30.0)
    ptr = C.cast(o.lib.orc_forward_offsets(), C.POINTER(C.c_int * 42))
    offs = {tuple(ptr.contents[3 * k:3 * k + 3]) for k in range(14)}
    neg = {(-a, -b, -c) for a, b, c in offs}
    assert len(offs) == 14 and offs & neg == {(0, 0, 0)} and len(offs | neg) == 27


def test_box_wrap_and_minimum_image_semantics():
    o = make(54.1)
    assert o.hinv[0, 0] == 1.0 / 54.1  # adjugate/determinant form happens to equal 1/L for this box
    p2 = make(64.0)  # power-of-two box: s = 0.5 exactly
    assert np.array_equal(p2.min_image([32.0, 0.0, 0.0]), [-32.0, 0.0, 0.0])       # round half away from zero
    assert np.array_equal(p2.min_image([-32.0, 0.0, 0.0]), [32.0, 0.0, 0.0])
    assert np.array_equal(p2.min_image([96.0, -31.0, 33.0]), [-32.0, -31.0, -31.0])
    w = p2.wrap([-1e-17, 64.0, 70.0])
    assert w[0] == 64.0 and w[1] == 0.0 and w[2] == 6.0                            # s - floor(s) may round to 1.0 -> x == L
    assert o.wrap([0.0, 54.1, 60.0])[1] == 54.1 * (54.1 * (1.0 / 54.1) - 0.0)      # L itself is (almost) a fixed point
    tri = Oracle([10.0, 0, 0, 2.0, 9.0, 0, 1.0, -1.5, 8.0])
    assert np.allclose(tri.h @ tri.hinv, np.eye(3), atol=1e-15)
    assert abs(tri.lib.orc_box_volume(tri.box) - 720.0) < 1e-12
    with pytest.raises(ValueError):
        Oracle([1.0, 0, 0, 2.0, 0, 0, 0, 0, 1.0])


def test_cell_binning_saturates_like_rust_casts():
    o = make(30.0, rc=9.0)
    n = o.divide_into_cells(9.0)
    assert n == (3, 3, 3)
    pos = np.array([[-0.5, 1.0, 1.0], [30.0, 1.0, 1.0], [31.0, 29.9, 1.0], [float("nan"), 1.0, 1.0]])
    cell_of, start, atoms = o.rcut_cells(pos, *n)
    assert cell_of[0] == 0          # negative -> saturates to cell 0 (not the periodic cell)
    assert cell_of[1] == 0          # s == 1.0 -> cell n -> % n == 0
    assert cell_of[2] == 0 + 2 * 3  # x wraps via %, y in the last cell
    assert cell_of[3] == 0          # NaN -> 0
    assert start[-1] == 4 and sorted(atoms.tolist()) == [0, 1, 2, 3]
    assert o.divide_into_cells(100.0) == (1, 1, 1)


def test_observables_and_step_against_numpy():
    atoms = fcc_argon(5, temperature=20.0, jitter=0.1, seed=2)
    L = atoms.sim_box.h[0, 0]
    o = make(L)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    _, f = o.compute_potential(x, atoms.type_ids)
    ke = o.kinetic_energy(v, atoms.type_ids)
    assert abs(ke - 0.5 * 39.948 * (v * v).sum()) < 1e-12 * ke
    assert abs(o.temperature(len(x), ke) - 2 * ke / (3 * len(x) * 0.0083144621)) < 1e-12
    assert abs(o.temperature(len(x), ke) - 20.0) < 1e-9  # create_velocities rescales to T exactly
    vir = o.virial_trace(x, f)
    assert abs(vir - (x * f).sum()) < 1e-10 * max(abs(vir), 1.0)
    assert abs(o.pressure(x, f, ke) - (2 * ke + vir) / (3 * L ** 3)) < 1e-15
    # one velocity-Verlet step, restated with numpy
    dt = 0.25
    a0 = f / 39.948
    x_np = x + (v * dt + (a0 * 0.5) * dt * dt)
    x_np -= np.floor(x_np / L) * L
    pe_np, f_np, _ = numpy_lj(x_np, L, EPS, SIG, 8.5)
    v_np = v + ((a0 + f_np / 39.948) * 0.5) * dt
    pe = o.verlet_step_nve(x, v, f, atoms.type_ids, dt)
    assert np.abs(x - x_np).max() < 1e-12 and np.abs(v - v_np).max() < 1e-13
    assert abs(pe - pe_np) < 1e-10 * abs(pe)


def test_nve_conserves_energy_and_serial_equals_omp():
    atoms = fcc_argon(5, temperature=5.0, seed=12345)
    L = atoms.sim_box.h[0, 0]
    o = make(L)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    th = o.run_nve(x, v, np.zeros_like(x), atoms.type_ids, 0.25, 200)
    h = th[1:, 2]
    assert np.abs(h - h[0]).max() < 2e-4 * abs(h[0])  # force is not shifted at rc: small drift is inherent
    x2, v2 = atoms.positions.copy(), atoms.velocities.copy()
    th2 = o.run_nve(x2, v2, np.zeros_like(x2), atoms.type_ids, 0.25, 200, mode="omp", threads=2)
    assert np.max(np.abs(th2[1:, 0] - th[1:, 0]) / np.abs(th[1:, 0])) < 1e-11


def test_missing_pair_and_key_order_quirk():
    """Table keys are stored as given and looked up sorted: an entry given as (2,1) is never found."""
    atoms = fcc_argon(4, temperature=0.0, jitter=0.1, seed=1)
    atoms.type_ids[::2] = 2
    L = atoms.sim_box.h[0, 0]
    good = Oracle.cubic(L, n_types=2, masses=(39.948, 20.0))
    bad = Oracle.cubic(L, n_types=2, masses=(39.948, 20.0))
    for o, key in ((good, (1, 2)), (bad, (2, 1))):
        o.insert(1, 1, EPS, SIG, 8.0)
        o.insert(2, 2, 0.1, 3.0, 7.0)
        o.insert(key[0], key[1], 0.15, 3.2, 7.5)
    pe_g, _ = good.compute_potential(atoms.positions, atoms.type_ids)
    pe_b, _ = bad.compute_potential(atoms.positions, atoms.type_ids)
    only_same = Oracle.cubic(L, n_types=2, masses=(39.948, 20.0))
    only_same.insert(1, 1, EPS, SIG, 8.0)
    only_same.insert(2, 2, 0.1, 3.0, 7.0)
    pe_s, _ = only_same.compute_potential(atoms.positions, atoms.type_ids)
    assert pe_g != pe_b
    # (2,1) is unreachable for pair lookup but still counts in max_rcut (cell grid) -> same energy as without it
    assert abs(pe_b - pe_s) < 1e-12 * abs(pe_s)
    assert bad.max_rcut() == 8.0 and good.max_rcut() == 8.0


def test_sqrt_free_threshold_is_exact():
    o = make(30.0)
    for rc in (8.5, 8.5125, 9.534, 2.5 * 3.405, 1.0, 7.0 / 3.0):
        t = o.rcut_threshold(rc)
        assert np.sqrt(t) <= rc < np.sqrt(np.nextafter(t, np.inf))


def test_golden_vectors():
    """tests/golden/oracle_small.json was produced by tests/golden/make_golden.py from this oracle; it pins
    the oracle against silent drift and is what the GPU tests also compare with."""
    with open(os.path.join(GOLDEN, "oracle_small.json")) as f:
        g = json.load(f)
    pos = np.array(g["positions"])
    vel = np.array(g["velocities"])
    types = np.array(g["types"], dtype=np.int32)
    o = make(g["L"], g["rc"])
    pe, f = o.compute_potential(pos, types)
    assert pe == g["pe"]
    assert np.array_equal(f, np.array(g["forces"]))
    start, nbr = o.build_neighbour_list(pos, types, extra=g["skin"])
    assert [sorted(nbr[start[i]:start[i + 1]].tolist()) for i in range(len(pos))] == g["neighbours_skin"]
    x, v, ff = pos.copy(), vel.copy(), np.zeros_like(pos)
    th = o.run_nve(x, v, ff, types, g["dt"], g["steps"])
    assert np.array_equal(th, np.array(g["thermo"]))
    assert np.array_equal(x, np.array(g["positions_end"]))


def test_oracle_matches_exact_arithmetic():
    """tests/golden/exact_small.json holds a 60-digit mpmath evaluation of the reference's FORMULAS (lennard_jones.rs:33-55,
    simulation_box.rs:17-42, potential.rs:15-33, properties.rs:17-65) on the committed 108-atom fixture, produced by
    tests/golden/make_exact.py without touching the oracle.  The oracle's f64 restatement must agree to 1e-13: this pins
    the oracle against the formulas themselves, not against its own output."""
    with open(os.path.join(GOLDEN, "oracle_small.json")) as f:
        g = json.load(f)
    with open(os.path.join(GOLDEN, "exact_small.json")) as f:
        e = json.load(f)
    # no in/out decision sits within rounding distance of a threshold: the exact sets are the f64 sets
    assert min(e["margin_rc0"], e["margin_list0"], e["margin_rc1"]) > 1e-9
    pos, vel, types = np.array(g["positions"]), np.array(g["velocities"]), np.array(g["types"], dtype=np.int32)
    o = make(g["L"], g["rc"])
    tol = 1e-13
    f0x = np.array(e["forces0"])
    fscale = np.sqrt((f0x ** 2).sum(axis=1).mean())
    for mode in ("serial", "omp", "n2"):
        pe, f = o.compute_potential(pos, types, mode=mode)
        assert abs(pe - e["pe0"]) <= tol * abs(e["pe0"]), mode
        assert np.abs(f - f0x).max() <= tol * fscale, mode
    start, nbr = o.build_neighbour_list(pos, types, extra=g["skin"])
    assert [sorted(nbr[start[i]:start[i + 1]].tolist()) for i in range(len(pos))] == e["neighbours_skin"]
    pe_l, f_l = o.compute_potential_list(pos, types, start, nbr)       # the list consumer (lennard_jones.rs:419-455)
    assert abs(pe_l - e["pe0"]) <= tol * abs(e["pe0"])
    assert np.abs(f_l - f0x).max() <= tol * fscale
    # one verlet_step_nve and the observables of the new state
    x, v, f = pos.copy(), vel.copy(), np.zeros_like(pos)
    th = o.run_nve(x, v, f, types, g["dt"], 1)
    assert abs(th[0, 0] - e["pe0"]) <= tol * abs(e["pe0"])
    assert np.abs(x - np.array(e["positions1"])).max() <= tol * g["L"]
    assert np.abs(v - np.array(e["velocities1"])).max() <= tol * np.abs(vel).max()
    assert np.abs(f - np.array(e["forces1"])).max() <= tol * fscale
    assert abs(th[1, 0] - e["pe1"]) <= tol * abs(e["pe1"])
    assert abs(th[1, 1] - e["ke1"]) <= tol * abs(e["ke1"])
    assert abs(th[1, 3] - e["temperature1"]) <= tol * abs(e["temperature1"])
    assert abs(o.virial_trace(x, f) - e["virial_trace1"]) <= 1e-12 * np.abs(x * f).sum()     # sum r.F cancels: bound by sum |r.F|
    assert abs(th[1, 4] - e["pressure1"]) <= 1e-12 * (2 * e["ke1"] + np.abs(x * f).sum()) / (3 * g["L"] ** 3)


def test_golden_ensemble_vectors():
    """tests/golden/oracle_small_ensembles.json pins the NVT / NPT restatements (nvt.rs, npt.rs, potential.rs:35-135)."""
    with open(os.path.join(GOLDEN, "oracle_small.json")) as f:
        g = json.load(f)
    with open(os.path.join(GOLDEN, "oracle_small_ensembles.json")) as f:
        e = json.load(f)
    pos, vel, types = np.array(g["positions"]), np.array(g["velocities"]), np.array(g["types"], dtype=np.int32)
    o = make(g["L"], g["rc"])
    chain = o.nhc_new(*e["temp"])
    th = o.run_nvt(pos.copy(), vel.copy(), np.zeros_like(pos), types, g["dt"], e["steps"], chain)
    assert np.array_equal(th, np.array(e["nvt_thermo"])) and list(chain.xi) == e["nvt_xi"]
    o = make(g["L"], g["rc"])
    chain = o.nhc_new(*e["temp"])
    baro = o.mtk_new(e["iso"][0], e["iso"][1], len(pos), e["temp"][0])
    x = pos.copy()
    th, htr = o.run_npt(x, vel.copy(), np.zeros_like(pos), types, g["dt"], e["steps"], baro, chain)
    assert np.array_equal(th, np.array(e["npt_thermo"])) and np.array_equal(htr, np.array(e["npt_h"]))
    assert list(baro.momentum) == e["npt_momentum"] and np.array_equal(x, np.array(e["npt_positions_end"]))
    assert abs(htr[-1][0] - htr[0][0]) > 1e-3      # the barostat moved the box in 25 steps


def test_mat3_exp_restatement_against_scipy():
    """nalgebra Matrix3::exp restated (Al-Mohy & Higham 2009) vs scipy.linalg.expm over every Pade branch."""
    import scipy.linalg as sl
    o = Oracle.cubic(30.0)
    rng = np.random.default_rng(5)
    for scale, tol in ((0.0, 0.0), (1e-9, 3e-16), (1e-5, 4e-16), (1e-2, 4e-16), (0.1, 1e-15), (0.6, 1e-13), (1.5, 1e-12), (6.0, 1e-10)):
        a = rng.normal(size=(3, 3)) * scale
        e = o.mat3_exp(a.T.reshape(9)).reshape(3, 3).T
        r = sl.expm(a)
        assert np.abs(e - r).max() <= max(tol, 1e-300) * max(np.abs(r).max(), 1.0)
    # symmetric generator (what MTKBarostat::scale feeds it): exp(A) exp(-A) = I
    a = rng.normal(size=(3, 3)) * 1e-3
    a = (a + a.T) * 0.5
    p = o.mat3_exp(a.T.reshape(9)).reshape(3, 3).T @ o.mat3_exp((-a).T.reshape(9)).reshape(3, 3).T
    assert np.abs(p - np.eye(3)).max() < 1e-15


def test_npt_restatement_properties():
    """verlet_step_npt_mtk restated (potential.rs:112-135): pressure tensor vs numpy, scale_box keeps fractional
    coordinates, zero barostat coupling (w -> inf) reduces to the NVT trace, and the box follows the sign of P - P_target."""
    atoms = fcc_argon(5, temperature=20.0, seed=11, jitter=0.1)
    o = Oracle.cubic(atoms.sim_box.h[0, 0])
    o.insert(1, 1, 0.238, 3.405, 8.5)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    pe, f = o.compute_potential(x, atoms.type_ids)
    pt = o.pressure_tensor(x, v, f).reshape(3, 3).T
    vol = o.box_volume()
    ref = (v.T @ v + x.T @ f) / vol
    assert np.abs(pt - ref).max() <= 1e-12 * np.abs(ref).max()
    # huge barostat mass: the box does not move and the trace equals NVT's
    steps = 20
    chain_a, chain_b = o.nhc_new(20.0, 30.0, 50.0), o.nhc_new(20.0, 30.0, 50.0)
    baro = o.mtk_new(0.0, 1e30, atoms.n_atoms, 20.0)
    xa, va = x.copy(), v.copy()
    tha, htr = o.run_npt(xa, va, np.zeros_like(x), atoms.type_ids, 0.25, steps, baro, chain_a)
    o2 = Oracle.cubic(atoms.sim_box.h[0, 0])
    o2.insert(1, 1, 0.238, 3.405, 8.5)
    xb, vb = x.copy(), v.copy()
    thb = o2.run_nvt(xb, vb, np.zeros_like(x), atoms.type_ids, 0.25, steps, chain_b)
    assert np.abs(htr - htr[0]).max() <= 1e-12
    assert np.abs(tha[:, :2] - thb[:, :2]).max() <= 1e-9 * np.abs(thb[:, :2]).max()
    # finite mass, target far above the instantaneous pressure: the box must shrink; fractional coordinates of a
    # pure scale_box are preserved
    o3 = Oracle.cubic(atoms.sim_box.h[0, 0])
    o3.insert(1, 1, 0.238, 3.405, 8.5)
    baro3 = o3.mtk_new(1.0, 50.0, atoms.n_atoms, 20.0)
    xc, vc = x.copy(), v.copy()
    thc, htc = o3.run_npt(xc, vc, np.zeros_like(x), atoms.type_ids, 0.25, 10, baro3, o3.nhc_new(20.0, 20.0, 50.0))
    assert htc[-1][0] < htc[0][0] and htc[-1][4] < htc[0][4] and htc[-1][8] < htc[0][8]
    s_before = x @ np.linalg.inv(htc[0].reshape(3, 3).T).T
    o4 = Oracle.cubic(atoms.sim_box.h[0, 0])
    xs = x.copy()
    scale = np.array([[1.01, 0.002, 0.0], [0.002, 0.99, 0.001], [0.0, 0.001, 1.003]])
    o4.scale_box(scale, xs)
    h_new = np.array(o4.box.h).reshape(3, 3).T
    assert np.abs(h_new - scale @ htc[0].reshape(3, 3).T).max() < 1e-12
    assert np.abs(xs @ np.linalg.inv(h_new).T - s_before).max() < 1e-12
