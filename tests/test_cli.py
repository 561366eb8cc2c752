"""The C++ host CLI (pis_b200/pis_b200_cli): input-script / LAMMPS-data readers (CPU, via --check), and the
full NVE run with thermo lines + dump.lammpstrj against the oracle (GPU).  The parser cases follow the
reference's own test module (src/tests/command_tests.rs), driven through the script instead of rstest."""
import json
import os
import subprocess
from decimal import Decimal

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "pis_b200", "pis_b200_cli")

DATA3 = """3 atoms
1 atom types

0.0 5.0 xlo xhi
0.0 5.0 ylo yhi
0.0 5.0 zlo zhi

Masses
1 39.948

Atoms
1 1 0.0 0.0 0.0   # trailing comments are tolerated (positional parsing)
2 1 1.0 1.0 1.0
3 1 2.5 2.5 2.5
"""


def rust_display(x: float) -> str:
    if x != x:
        return "NaN"
    s = format(Decimal(repr(float(x))), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s if s not in ("", "-") else "0"


def check(tmp_path, script: str, data: str | None = None, expect_fail=False):
    if data is not None:
        (tmp_path / "data.txt").write_text(data)
    (tmp_path / "in.pis").write_text(script)
    r = subprocess.run([CLI, "-i", "in.pis", "--check"], cwd=tmp_path, capture_output=True, text=True)
    if expect_fail:
        assert r.returncode == 1 and r.stderr.startswith("Error: "), (r.returncode, r.stderr)
        return r.stderr.strip()
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout)


def test_defaults_and_basic_commands(tmp_path):
    ctx = check(tmp_path, "timestep 0.005\nrun 250\n")
    assert ctx["timestep"] == 0.005 and ctx["steps"] == 250 and ctx["ensemble"] == "NVE"
    assert ctx["dump"] == {"name": "default_dump", "group": "all", "style": "atoms", "dump_step": 1, "file_name": "dump.lammpstrj"}
    ctx = check(tmp_path, "# comment\n\nrun\n   timestep 0.25 # inline comment\n")  # <2 tokens -> ignored
    assert ctx["timestep"] == 0.25 and ctx["steps"] == 100
    ctx = check(tmp_path, "dump dp1 all atom 10 dump.lammpstrj\n")
    assert ctx["dump"]["dump_step"] == 10 and ctx["dump"]["style"] == "atom" and ctx["dump"]["name"] == "dp1"


@pytest.mark.parametrize("script,needle", [
    ("frobnicate 1 2\n", "Invalid command frobnicate found line: 1"),
    ("timestep abc\n", "Error parsing floating number from string abc"),
    ("run -5\n", "Negative value -5 not allowed on line: 1"),
    ("run 1.5\n", "Error parsing integer number from string 1.5: invalid digit found in string"),
    ("run 99999999999\n", "number too large to fit in target type"),
    ("velocity all set 1 2 3\n", "Invalid argument: set at line: 1"),
    ("velocity all create 5.0 1 dist triangular\n", "Invalid argument: triangular"),
    ("velocity all create 5.0 1 loop geom\n", "Invalid argument: loop"),
    ("velocity all\n", "Missing argument on line 1"),
    ("pair_coeff 1 1 0.2 3.4\n", "Potential manager not initialized"),
    ("dump d all atom x out.traj\n", "Error parsing integer number from string x"),
    ("dump d all atom 5\n", "Missing argument on line 1"),
    ("fix f all nvt temp 1.0 2.0\n", "Missing argument on line 1"),
    ("fix f all nvt pressure 1 2 3\n", "Invalid argument: pressure"),
    ("read_data nope.txt\n", "Failed to open input file 'nope.txt'"),
])
def test_parser_errors(tmp_path, script, needle):
    assert needle in check(tmp_path, script, expect_fail=True)


@pytest.mark.parametrize("script,needle", [
    # src/tests/command_tests.rs, re-expressed as script lines (its `line: 0` becomes the script's line number)
    ("TIMESTEP 0.1\n", "Invalid command TIMESTEP found line: 1"),            # from_str_unknown_command_returns_none: keywords are case-sensitive
    ("TimeStep 0.1\n", "Invalid command TimeStep found line: 1"),
    ("timestep notanumber\n", "Error parsing floating number from string notanumber"),   # timestep_invalid_float_returns_float_parse_error
    ("timestep one\n", "Error parsing floating number from string one"),
    ("timestep 0.1.2\n", "Error parsing floating number from string 0.1.2"),
    ("run -1\n", "Negative value -1 not allowed on line: 1"),                 # runsteps_negative_value_returns_negative_value_error
    ("run ten\n", "Error parsing integer number from string ten"),            # runsteps_non_integer_returns_int_parse_error
    ("run 1e3\n", "Error parsing integer number from string 1e3"),
    ("velocity all scale 300.0 42\n", "Invalid argument: scale at line: 1"),  # velocity_unknown_style_returns_invalid_argument
    ("velocity all create 300.0 42 badkeyword value\n", "Invalid argument: badkeyword"),  # velocity_invalid_keyword_or_dist_...
    ("velocity all create 300.0 42 dist lorentzian\n", "Invalid argument: lorentzian"),
    ("dump my_dump all atom 100\n", "Missing argument on line 1"),            # dump_missing_args_returns_missing_argument
    ("dump my_dump all atom\n", "Missing argument on line 1"),
    ("dump my_dump all\n", "Missing argument on line 1"),
    ("dump my_dump all atom every output.lammpstrj\n", "Error parsing integer number from string every"),   # dump_non_integer_step_...
    ("dump my_dump all atom 1.5 output.lammpstrj\n", "Error parsing integer number from string 1.5"),
    ("fix myfix all\n", "Missing argument on line 1"),                        # fix_missing_required_args_returns_missing_argument
    ("fix myfix all nvt badkeyword 300.0 300.0 0.1\n", "Invalid argument: badkeyword"),   # fix_unknown_keyword_returns_invalid_argument
    ("fix myfix all npt unknown 1.0 1.0 1.0\n", "Invalid argument: unknown"),
    ("\n\n# c\nrun 5\nfix myfix all npt unknown 1 1 1\n", "at line: 5"),    # errors name the script line
])
def test_parser_errors_from_the_reference_test_module(tmp_path, script, needle):
    assert needle in check(tmp_path, script, expect_fail=True)


def test_statements_from_the_reference_test_module(tmp_path):
    """The success cases of src/tests/command_tests.rs through the script reader."""
    assert check(tmp_path, "timestep 0.001\n")["timestep"] == 0.001                                  # timestep_sets_value_on_ctx
    assert check(tmp_path, "run 5000\n")["steps"] == 5000                                            # runsteps_sets_value_on_ctx
    v = check(tmp_path, "velocity all create 300.0 42\n")["velocity"]                                # velocity_create_minimal_sets_ctx
    assert v["group"] == "all" and v["temperature"] == 300 and v["seed"] == 42 and "dist" not in v
    assert check(tmp_path, "velocity all create 300.0\n")["velocity"]["seed"] == 0                   # ..._without_seed_ctx
    for dist in ("gaussian", "uniform"):                                                             # velocity_create_dist_keyword_sets_distribution
        assert check(tmp_path, f"velocity all create 300.0 42 dist {dist}\n")["velocity"]["dist"] == dist
    v = check(tmp_path, "velocity all create 300.0 dist gaussian\n")["velocity"]                     # ..._without_seed
    assert v["seed"] == 0 and v["dist"] == "gaussian"
    d = check(tmp_path, "dump my_dump all atom 100 output.lammpstrj\n")["dump"]                      # dump_sets_all_fields_on_ctx
    assert d == {"name": "my_dump", "group": "all", "style": "atom", "dump_step": 100, "file_name": "output.lammpstrj"}
    # pair_style stores its words verbatim and pair_coeff appends to it (pair_style_stores_args_in_ctx, pair_coeff_appends_...)
    p = check(tmp_path, "pair_style lj/cut 10.0\npair_coeff 1 1 0.238 3.405\npair_coeff 1 2 0.1 3.0 9.0\n")["potential"]
    assert p["max_rcut"] == 10 and [(q["i"], q["j"], q["rcut"]) for q in p["pairs"]] == [(1, 1, 10), (1, 2, 9)]
    # fix_nvt_temp_keyword_sets_nh_chain_args / fix_npt_sets_both_barostat_and_thermostat
    assert check(tmp_path, "fix myfix all nvt temp 300.0 300.0 0.1\n")["ensemble"] == "NVT"
    assert check(tmp_path, "fix myfix all npt temp 300.0 300.0 0.1 iso 1.0 1.0 1.0\n")["ensemble"] == "NPT"


def test_two_type_data_file(tmp_path):
    """The shape of the reference's example/data_run.txt: two atom types, two Masses rows, one PairCoeffs row per type."""
    data = ("4 atoms\n2 atom types\n\n0.0 12.0 xlo xhi\n0.0 12.0 ylo yhi\n0.0 12.0 zlo zhi\n\nMasses\n1 39.948\n2 20.18\n\n"
            "PairCoeffs\n1 0.238 3.405 8.5\n2 0.07 2.8\n\nAtoms\n1 1 1.0 1.0 1.0\n2 2 4.0 4.0 4.0\n4 2 7.0 7.5 8.0\n3 1 9.0 9.0 9.0\n")
    ctx = check(tmp_path, "read_data data.txt\n", data)
    a = ctx["atoms"]
    assert a["n_atoms"] == 4 and a["n_types"] == 2 and a["masses"] == [39.948, 20.18] and a["types_head"] == [1, 2, 1, 2]
    assert a["first_positions"][6:12] == [9, 9, 9, 7, 7.5, 8]                 # rows land at their ids, in any order
    assert ctx["potential"]["pairs"] == [{"i": 1, "j": 1, "epsilon": 0.238, "sigma": 3.405, "rcut": 8.5},
                                         {"i": 2, "j": 2, "epsilon": 0.07, "sigma": 2.8, "rcut": 7}]
    assert "Atom type 3 out of range" in check(tmp_path, "read_data data.txt\n", data.replace("2 20.18", "3 20.18"), expect_fail=True)


def test_velocity_command(tmp_path):
    ctx = check(tmp_path, "velocity all create 300.0 12345 dist gaussian\n")
    assert ctx["velocity"] == {"group": "all", "start_velocity": True, "temperature": 300, "seed": 12345, "dist": "gaussian"}
    ctx = check(tmp_path, "velocity all create 10.0\n")   # missing seed -> 0
    assert ctx["velocity"]["seed"] == 0
    ctx = check(tmp_path, "velocity all create 10.0 dist uniform\n")  # non-integer seed token -> seed 0, token re-read as keyword
    assert ctx["velocity"]["seed"] == 0 and ctx["velocity"]["dist"] == "uniform"


def test_fix_selects_ensemble(tmp_path):
    assert check(tmp_path, "fix a all nvt temp 5.0 5.0 50\n")["ensemble"] == "NVT"
    assert check(tmp_path, "fix a all npt temp 5.0 50.0 100 iso 0.01 0.01 1000\n")["ensemble"] == "NPT"
    assert check(tmp_path, "fix a all nve temp 5.0 50.0 100\n")["ensemble"] == "NVE"   # unknown style: silently ignored
    assert check(tmp_path, "fix a all nvt iso 1 1 1\n")["ensemble"] == "NVE"           # iso without npt: ignored


def test_read_data(tmp_path):
    ctx = check(tmp_path, "read_data data.txt\n", DATA3)
    a = ctx["atoms"]
    assert a["n_atoms"] == 3 and a["box"] == [5, 5, 5] and a["masses"] == [39.948]
    assert a["first_positions"] == [0, 0, 0, 1, 1, 1, 2.5, 2.5, 2.5]
    assert ctx["potential"] is None
    # with neither a Velocities section nor a velocity command, 300 K / seed 0 velocities are created silently
    assert ctx["velocity"]["start_velocity"] is True and any(v != 0 for v in a["first_velocities"])
    for bad_id in ("0", "4"):
        msg = check(tmp_path, "read_data data.txt\n", DATA3.replace("3 1 2.5", f"{bad_id} 1 2.5"), expect_fail=True)
        assert f"Atom count mismatch: expected 3, found {bad_id}" in msg


def test_velocities_section_and_paircoeffs(tmp_path):
    data = DATA3 + "\nVelocities\n1 0.1 0.2 0.3\n2 0.0 0.0 0.0\n3 -0.1 -0.2 -0.3\n\nPairCoeffs\n1 0.238 3.405 8.5\n"
    ctx = check(tmp_path, "read_data data.txt\n", data)
    assert ctx["velocity"]["start_velocity"] is False
    assert ctx["atoms"]["first_velocities"] == [0.1, 0.2, 0.3, 0, 0, 0, -0.1, -0.2, -0.3]
    assert ctx["potential"]["pairs"] == [{"i": 1, "j": 1, "epsilon": 0.238, "sigma": 3.405, "rcut": 8.5}]
    # a velocity command AFTER read_data replaces the context and re-randomises (commands.rs:80,141)
    ctx = check(tmp_path, "read_data data.txt\nvelocity all create 5.0 7\n", data)
    assert ctx["velocity"]["start_velocity"] is True
    # default rc = 2.5 sigma; and the reference's "i j eps sigma rc" mis-parse (token 1 read as epsilon)
    ctx = check(tmp_path, "read_data data.txt\n", DATA3 + "\nPairCoeffs\n1 0.5 2.0\n")
    assert ctx["potential"]["pairs"][0]["rcut"] == 5
    ctx = check(tmp_path, "read_data data.txt\n", DATA3.replace("1 atom types", "2 atom types") + "\nPairCoeffs\n1 2 0.6 1.1 2.8\n")
    assert ctx["potential"]["pairs"] == [{"i": 1, "j": 1, "epsilon": 2, "sigma": 0.6, "rcut": 1.1}]


def test_pair_style_replaces_data_file_potential(tmp_path):
    data = DATA3 + "\nPairCoeffs\n1 0.238 3.405 8.5\n"
    ctx = check(tmp_path, "read_data data.txt\npair_style lj/cut 7.0\npair_coeff 1 1 0.1 3.0\npair_coeff 2 1 0.2 3.1 6.5\n", data)
    assert ctx["potential"]["pairs"] == [{"i": 1, "j": 1, "epsilon": 0.1, "sigma": 3, "rcut": 7},
                                         {"i": 2, "j": 1, "epsilon": 0.2, "sigma": 3.1, "rcut": 6.5}]   # key kept as given
    assert "Unknown pair style: 'eam'" in check(tmp_path, "pair_style eam 1.0\n", expect_fail=True)


def test_cli_errors(tmp_path):
    r = subprocess.run([CLI, "-i", "missing.pis"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to open input file 'missing.pis'" in r.stderr
    (tmp_path / "in.pis").write_text("read_data data.txt\nrun 3\n")
    (tmp_path / "data.txt").write_text(DATA3)
    r = subprocess.run([CLI, "-i", "in.pis"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "Potential manager not initialized" in r.stderr


@pytest.mark.gpu
def test_cli_nvt_run_matches_oracle(tmp_path):
    """`fix ... nvt temp 5.0 50.0 50` through the CLI: thermo lines (H includes the thermostat energy) vs the oracle."""
    from oracle.pis_oracle import Oracle
    from pis_b200.lattice import fcc_argon

    atoms = fcc_argon(6, temperature=5.0, seed=3)
    n, L = atoms.n_atoms, atoms.sim_box.h[0, 0]
    lines = [f"{n} atoms", "1 atom types", "", f"0.0 {rust_display(L)} xlo xhi", f"0.0 {rust_display(L)} ylo yhi",
             f"0.0 {rust_display(L)} zlo zhi", "", "Masses", "1 39.948", "", "PairCoeffs", "1 0.238 3.405 8.5", "", "Atoms"]
    lines += [f"{i + 1} 1 {repr(float(p[0]))} {repr(float(p[1]))} {repr(float(p[2]))}" for i, p in enumerate(atoms.positions)]
    lines += ["", "Velocities"]
    lines += [f"{i + 1} {repr(float(v[0]))} {repr(float(v[1]))} {repr(float(v[2]))}" for i, v in enumerate(atoms.velocities)]
    (tmp_path / "argon.txt").write_text("\n".join(lines) + "\n")
    steps = 60
    (tmp_path / "input.pis").write_text(f"timestep 0.25\nread_data argon.txt\nfix mynvt all nvt temp 5.0 50.0 50\n"
                                        f"dump dp1 all atom 20 dump.lammpstrj\nrun {steps}\n")
    r = subprocess.run([CLI, "-i", "input.pis", "--skin", "1.0215"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout.strip().splitlines()
    o = Oracle.cubic(L)
    o.insert(1, 1, 0.238, 3.405, 8.5)
    chain = o.nhc_new(5.0, 50.0, 50.0)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref = o.run_nvt(x, v, np.zeros_like(x), atoms.type_ids, 0.25, steps, chain)
    assert len(out) == steps + 1
    for s in range(1, steps + 1):
        got = [float(t) for t in out[s].split()[1:]]
        for g, e in zip(got, ref[s]):
            assert abs(g - e) <= 1.1e-3 + 1e-9 * abs(e)


@pytest.mark.gpu
def test_cli_npt_run_matches_oracle(tmp_path):
    """The reference's shipped example/input.pis line `fix mynpt all npt temp 5.0 50.0 100 iso 0.01 0.01 1000` through the
    CLI: thermo lines (H includes thermostat + barostat energy, P uses the step's volume) and the dump's changing box bounds."""
    from oracle.pis_oracle import Oracle
    from pis_b200.lattice import fcc_argon

    atoms = fcc_argon(6, temperature=5.0, seed=3)
    n, L = atoms.n_atoms, atoms.sim_box.h[0, 0]
    lines = [f"{n} atoms", "1 atom types", "", f"0.0 {rust_display(L)} xlo xhi", f"0.0 {rust_display(L)} ylo yhi",
             f"0.0 {rust_display(L)} zlo zhi", "", "Masses", "1 39.948", "", "PairCoeffs", "1 0.238 3.405 8.5", "", "Atoms"]
    lines += [f"{i + 1} 1 {repr(float(p[0]))} {repr(float(p[1]))} {repr(float(p[2]))}" for i, p in enumerate(atoms.positions)]
    lines += ["", "Velocities"]
    lines += [f"{i + 1} {repr(float(v[0]))} {repr(float(v[1]))} {repr(float(v[2]))}" for i, v in enumerate(atoms.velocities)]
    (tmp_path / "argon.txt").write_text("\n".join(lines) + "\n")
    steps = 60
    (tmp_path / "input.pis").write_text(f"timestep 0.25\nread_data argon.txt\nfix mynpt all npt temp 5.0 50.0 100 iso 0.01 0.01 1000\n"
                                        f"dump dp1 all atom 20 dump.lammpstrj\nrun {steps}\n")
    r = subprocess.run([CLI, "-i", "input.pis", "--skin", "1.0215"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout.strip().splitlines()
    o = Oracle.cubic(L)
    o.insert(1, 1, 0.238, 3.405, 8.5)
    chain = o.nhc_new(5.0, 50.0, 100.0)
    baro = o.mtk_new(0.01, 1000.0, n, 5.0)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    ref, htr = o.run_npt(x, v, np.zeros_like(x), atoms.type_ids, 0.25, steps, baro, chain)
    assert len(out) == steps + 1
    for s in range(1, steps + 1):
        got = [float(t) for t in out[s].split()[1:]]
        for g, e in zip(got, ref[s]):
            assert abs(g - e) <= 1.1e-3 + 1e-9 * abs(e)
    dump = (tmp_path / "dump.lammpstrj").read_text().splitlines()
    frame = 9 + n
    assert len(dump) == frame * (steps // 20 + 1)
    for k in range(steps // 20 + 1):
        blk = dump[k * frame:(k + 1) * frame]
        hs = htr[20 * k].reshape(3, 3).T
        for d in range(3):  # bounds are `0 h_dd` from the diagonal of the step's box (dump_traj.rs:44-50)
            assert abs(float(blk[5 + d].split()[1]) - hs[d, d]) <= 1e-9 * hs[d, d]
    rows = np.array([[float(t) for t in ln.split()] for ln in dump[-n:]])
    assert np.abs(rows[:, 2:] - x).max() < 1e-8
    assert abs(htr[-1][0] - L) > 1e-6  # the barostat did move the box


@pytest.mark.gpu
def test_cli_nve_run_matches_oracle(tmp_path):
    """example/input.pis shape (NVE): thermo lines and dump.lammpstrj against the oracle's Simulation::run."""
    from oracle.pis_oracle import Oracle
    from pis_b200.lattice import fcc_argon

    atoms = fcc_argon(6, temperature=5.0, seed=12345)
    n, L = atoms.n_atoms, atoms.sim_box.h[0, 0]
    lines = [f"{n} atoms", "1 atom types", "", f"0.0 {rust_display(L)} xlo xhi", f"0.0 {rust_display(L)} ylo yhi",
             f"0.0 {rust_display(L)} zlo zhi", "", "Masses", "1 39.948", "", "PairCoeffs", "1 0.238 3.405 8.5", "", "Atoms"]
    lines += [f"{i + 1} 1 {repr(float(p[0]))} {repr(float(p[1]))} {repr(float(p[2]))}" for i, p in enumerate(atoms.positions)]
    lines += ["", "Velocities"]
    lines += [f"{i + 1} {repr(float(v[0]))} {repr(float(v[1]))} {repr(float(v[2]))}" for i, v in enumerate(atoms.velocities)]
    (tmp_path / "argon.txt").write_text("\n".join(lines) + "\n")
    steps = 40
    (tmp_path / "input.pis").write_text(f"timestep 0.25\nread_data argon.txt\ndump dp1 all atom 10 dump.lammpstrj\nrun {steps}\n")
    r = subprocess.run([CLI, "-i", "input.pis", "--skin", "1.0215"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout.strip().splitlines()

    o = Oracle.cubic(L)
    o.insert(1, 1, 0.238, 3.405, 8.5)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    snaps = {0: x.copy()}
    f = np.zeros_like(x)
    pe0, _ = o.compute_potential(x, atoms.type_ids, forces=f)
    assert out[0].split()[0] == "0" and abs(float(out[0].split()[1]) - pe0) <= 1e-9 * abs(pe0)
    assert len(out) == steps + 1
    for s in range(1, steps + 1):
        pe = o.verlet_step_nve(x, v, f, atoms.type_ids, 0.25)
        ke = o.kinetic_energy(v, atoms.type_ids)
        ref = [pe, ke, pe + ke, o.temperature(n, ke), o.pressure(x, f, ke)]
        got = out[s].split()
        assert got[0] == str(s) and len(got) == 6
        for g, e in zip(got[1:], ref):
            assert len(g.split(".")[1]) == 3                      # {:.3}
            assert abs(float(g) - e) <= 1.1e-3 + 1e-9 * abs(e)
        if s % 10 == 0:
            snaps[s] = x.copy()
    dump = (tmp_path / "dump.lammpstrj").read_text().splitlines()
    frame = 9 + n
    assert len(dump) == frame * (steps // 10 + 1)
    for k, s in enumerate(sorted(snaps)):
        blk = dump[k * frame:(k + 1) * frame]
        assert blk[:9] == ["ITEM: TIMESTEP", str(s), "ITEM: NUMBER OF ATOMS", str(n), "ITEM: BOX BOUNDS pp pp pp",
                           f"0 {rust_display(L)}", f"0 {rust_display(L)}", f"0 {rust_display(L)}", "ITEM: ATOMS id type x y z"]
        rows = np.array([[float(t) for t in ln.split()] for ln in blk[9:]])
        assert np.array_equal(rows[:, 0], np.arange(1, n + 1)) and (rows[:, 1] == 1).all()
        assert np.abs(rows[:, 2:] - snaps[s]).max() < 1e-9
        if s == 0:  # frame 0 is the input, printed with Rust `{}` float formatting
            assert blk[9 + 1] == "2 1 " + " ".join(rust_display(c) for c in atoms.positions[1])


@pytest.mark.gpu
def test_cli_velocity_create_runs_on_the_device(tmp_path):
    """`velocity all create T seed` after read_data (no Velocities section): the CLI generates the velocities on the device
    (pisb_start_velocities) once the atoms are resident.  The thermo trace must be the oracle's run from the HOST
    generator's velocities for the same (T, seed) -- the two generators are the same function of (seed, atom id) -- and
    step 1 must show T close to the requested 20 K (rescale_to_temperature, velocities.rs:52-59)."""
    from oracle.pis_oracle import Oracle
    from pis_b200.lattice import create_velocities, fcc_argon

    atoms = fcc_argon(6, temperature=0.0, seed=5, jitter=0.05)
    n, L = atoms.n_atoms, atoms.sim_box.h[0, 0]
    lines = [f"{n} atoms", "1 atom types", "", f"0.0 {rust_display(L)} xlo xhi", f"0.0 {rust_display(L)} ylo yhi",
             f"0.0 {rust_display(L)} zlo zhi", "", "Masses", "1 39.948", "", "PairCoeffs", "1 0.238 3.405 8.5", "", "Atoms"]
    lines += [f"{i + 1} 1 {repr(float(p[0]))} {repr(float(p[1]))} {repr(float(p[2]))}" for i, p in enumerate(atoms.positions)]
    (tmp_path / "argon.txt").write_text("\n".join(lines) + "\n")
    steps = 30
    (tmp_path / "input.pis").write_text(f"timestep 0.25\nread_data argon.txt\nvelocity all create 20.0 77\nrun {steps}\n")
    r = subprocess.run([CLI, "-i", "input.pis", "--skin", "1.0215"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout.strip().splitlines()
    assert len(out) == steps + 1

    o = Oracle.cubic(L)
    o.insert(1, 1, 0.238, 3.405, 8.5)
    x = atoms.positions.copy()
    v = create_velocities(n, np.full(n, 39.948), 20.0, 77)
    f = np.zeros_like(x)
    o.compute_potential(x, atoms.type_ids, forces=f)
    for s in range(1, steps + 1):
        pe = o.verlet_step_nve(x, v, f, atoms.type_ids, 0.25)
        ke = o.kinetic_energy(v, atoms.type_ids)
        ref = [pe, ke, pe + ke, o.temperature(n, ke), o.pressure(x, f, ke)]
        got = out[s].split()
        for g, e in zip(got[1:], ref):
            assert abs(float(g) - e) <= 1.1e-3 + 1e-9 * abs(e)
    assert abs(float(out[1].split()[4]) - 20.0) < 0.5
