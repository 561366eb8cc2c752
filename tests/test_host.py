"""CPU tests of the host-side mirror of the reference interface (no GPU, no oracle in the product path)."""
import json
import os

import numpy as np
import pytest

from pis_b200 import Atoms, LennardJones, LJCudaManager, SimulationBox
from pis_b200.lattice import ARGON, fcc_argon, fcc_positions

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_simulation_box_inverse_and_volume():
    b = SimulationBox.from_lammps_data(0.0, 54.1, 0.0, 54.1, 0.0, 54.1)
    assert b.h_inv[0, 0] == (54.1 * 54.1) / (54.1 * (54.1 * 54.1))
    assert b.pbc == [True, True, True] and abs(b.volume() - 54.1 ** 3) < 1e-9
    t = SimulationBox.from_lammps_data(1.0, 11.0, -2.0, 7.0, 0.0, 8.0, xy=2.0, xz=1.0, yz=-1.5)
    assert np.allclose(t.h, [[10, 2, 1], [0, 9, -1.5], [0, 0, 8]])
    assert np.allclose(t.h_inv, np.linalg.inv(t.h), rtol=1e-14, atol=1e-16)
    assert np.array_equal(t.h_colmajor(), t.h.T.reshape(9))
    with pytest.raises(ValueError):
        SimulationBox([[1, 2, 0], [0, 0, 0], [0, 0, 1]])


def test_host_box_matches_oracle_box_bitwise():
    """Two independent statements of nalgebra's try_inverse (product host code vs test oracle)."""
    from oracle.pis_oracle import Oracle

    for h in ([[54.1, 0, 0], [0, 54.1, 0], [0, 0, 54.1]], [[10, 2, 1], [0, 9, -1.5], [0, 0, 8]],
              [[541.0, 0, 0], [0, 270.5, 0], [0, 0, 1082.0]]):
        b = SimulationBox(h)
        o = Oracle(b.h_colmajor())
        assert np.array_equal(o.hinv, b.h_inv)


def test_atoms_scalar_observables():
    a = fcc_argon(3, temperature=12.0)
    ke = float((0.5 * 39.948 * (a.velocities ** 2).sum()))
    assert abs(a.temerature(ke) - 12.0) < 1e-9
    assert a.degress_of_freedom() == 3 * 108 and a.mass_i(5) == 39.948
    assert abs(a.pressure(ke, -3.0) - (2 * ke - 3.0) / (3 * a.sim_box.volume())) < 1e-18
    assert a.positions.flags.c_contiguous and a.positions.shape == (108, 3)
    assert np.abs((a.velocities * 39.948).sum(axis=0)).max() < 1e-10  # drift removed


def test_fcc_generator_order_matches_reference_example():
    """example/argon4000.txt was written by the reference's sapphire generator: same order, same values."""
    with open(os.path.join(GOLDEN, "argon4000_head.json")) as f:
        lines = json.load(f)["lines"]
    assert ["4000", "atoms"] in [ln.split()[:2] for ln in lines]
    assert "0.0 54.1 xlo xhi" in lines and "1 0.238 3.405 8.5" in lines
    rows = [ln.split() for ln in lines if len(ln.split()) == 5 and ln.split()[0].isdigit()]
    pos = fcc_positions(ARGON["a"], 10, 10, 10)
    assert len(rows) >= 12
    for r in rows:
        i = int(r[0]) - 1
        assert r[1] == "1"
        assert [float(v) for v in r[2:]] == pos[i].tolist()  # bit-identical to the printed shortest repr


def test_manager_table_semantics_without_gpu():
    m = LJCudaManager.new(skin=1.0)
    assert m.is_empty() and m.max_rcut() == 0.0
    m.insert((2, 1), LennardJones(0.1, 3.0, 9.0))
    m.insert((1, 1), LennardJones(0.238, 3.405, 8.5))
    assert not m.is_empty() and m.max_rcut() == 9.0
    a = fcc_argon(2, temperature=0.0)
    a.type_ids[1] = 2
    a.masses = [39.948, 20.0]
    assert m.get_potential_ij(a, 0, 1) is None          # (1,2) was never inserted: key order quirk
    assert m.get_potential_ij(a, 0, 2).get_rcut() == 8.5
    assert m.get((2, 1)).sigma == 3.0


def test_product_fails_loudly_without_gpu():
    from pis_b200 import capi

    if capi.load().pisb_device_count() > 0:
        pytest.skip("a GPU is present")
    m = LJCudaManager(skin=1.0)
    m.insert((1, 1), LennardJones(0.238, 3.405, 8.5))
    with pytest.raises(capi.PisbError) as e:
        m.compute_potential(fcc_argon(3, temperature=0.0))
    assert e.value.code == capi.PISB_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for base, _dirs, files in os.walk(os.path.join(root, "pis_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "pis_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_roofline_traffic_file_matches_the_kernel_the_bench_reports():
    """bench.py takes roofline.traffic (ncu dram bytes per launch) from profiles/force_traffic.json only when that capture is
    of the kernel it times: the default 4M-atom run steps with k_force_vv."""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "force_traffic.json")) as f:
        t = json.load(f)
    assert t["kernel"].startswith("k_force_vv")
    assert 1.5e9 < t["dram_bytes_per_launch"] < 4e9 and 0 < t["fp64_pipe_active_pct"] <= 100
    # algorithmic bytes of the same launch: (192 + 4K) per atom at K ~ 85, 4M atoms
    assert 0.9 < t["dram_bytes_per_launch"] / ((192 + 4 * 85.3) * 4.0e6) < 1.3


def test_every_library_option_is_documented_in_the_header():
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "pis_b200", "csrc", "pisb_sim.cu")).read()
    body = src[src.index("int pisb_set_option("):]
    names = set(re.findall(r'std::strcmp\(name, "(\w+)"\)', body))
    header = open(os.path.join(root, "include", "pisb200.h")).read()
    assert names and all(n in header for n in names), sorted(n for n in names if n not in header)


def test_example_inputs_are_the_reference_bytes():
    """BASELINE configs[0] travels to the GPU box as tests/golden/example_input.pis (byte copy of the reference's
    example/input.pis) and a regenerated example/argon4000.txt; both must hash to the reference's files (sha256 recorded by
    tests/golden/make_example_inputs.py), and to the files themselves where /root/reference exists."""
    import hashlib
    import json

    from tests.example_inputs import GOLDEN, argon4000_text, example_script

    with open(os.path.join(GOLDEN, "example_sha256.json")) as f:
        sha = json.load(f)
    with open(os.path.join(GOLDEN, "example_input.pis"), "rb") as f:
        script = f.read()
    data = argon4000_text().encode()
    assert hashlib.sha256(script).hexdigest() == sha["example/input.pis"]
    assert hashlib.sha256(data).hexdigest() == sha["example/argon4000.txt"]
    ref = "/root/reference/example"
    if os.path.isdir(ref):
        assert open(os.path.join(ref, "input.pis"), "rb").read() == script
        assert open(os.path.join(ref, "argon4000.txt"), "rb").read() == data
    nve = example_script(nve=True, steps=7)
    assert "npt" not in [t for ln in nve.split("\n") for t in ln.split("#")[0].split()] and "run 7" in nve
    assert "fix mynpt all npt temp 5.0 50.0 100 iso 0.01 0.01 1000" in example_script()
