"""Generate the committed golden fixtures.

  python tests/golden/make_golden.py

* oracle_small.json   : a 108-atom jittered FCC-argon system -- positions, velocities, step-0 PE and forces,
                        the full neighbour list with skin, and a 25-step NVE thermo trace, all produced by the
                        CPU oracle (oracle/pis_oracle.c).  NOT reference output: the Rust reference cannot be
                        built in this image (parity unpinned).
* oracle_small_ensembles.json : 25-step NVT (Nose-Hoover chain) and NPT (MTK barostat) traces of the same system from the
                        oracle's restatement of nvt.rs / npt.rs / potential.rs:35-135 (thermo rows, chain, box and barostat).
* argon4000_head.json : the header and first 16 atoms of the reference's example/argon4000.txt, read from
                        /root/reference when present -- pins the FCC generator's atom order and the reader.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.pis_oracle import Oracle  # noqa: E402
from pis_b200.lattice import fcc_argon  # noqa: E402


def main():
    atoms = fcc_argon(3, temperature=30.0, seed=21, jitter=0.12)
    L = float(atoms.sim_box.h[0, 0])
    rc, skin, dt, steps = 4.6, 0.8, 0.25, 25   # L = 16.23 >= 3 * (rc + skin): 3 cells per edge (n < 3 double-counts in the reference)
    o = Oracle.cubic(L)
    o.insert(1, 1, 0.238, 3.405, rc)
    pe, f = o.compute_potential(atoms.positions, atoms.type_ids)
    start, nbr = o.build_neighbour_list(atoms.positions, atoms.type_ids, extra=skin)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    th = o.run_nve(x, v, np.zeros_like(x), atoms.type_ids, dt, steps)
    g = {"L": L, "rc": rc, "skin": skin, "dt": dt, "steps": steps, "eps": 0.238, "sigma": 3.405, "mass": 39.948,
         "positions": atoms.positions.tolist(), "velocities": atoms.velocities.tolist(),
         "types": atoms.type_ids.tolist(), "pe": pe, "forces": f.tolist(),
         "neighbours_skin": [sorted(nbr[start[i]:start[i + 1]].tolist()) for i in range(atoms.n_atoms)],
         "thermo": th.tolist(), "positions_end": x.tolist()}
    with open(os.path.join(HERE, "oracle_small.json"), "w") as fh:
        json.dump(g, fh)
    # ensemble wrappers on the same system: `fix nvt temp 30 60 20` and `fix npt temp 30 60 20 iso 0.02 0.02 60`
    ens = {"temp": [30.0, 60.0, 20.0], "iso": [0.02, 60.0], "steps": steps}
    o2 = Oracle.cubic(L)
    o2.insert(1, 1, 0.238, 3.405, rc)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    chain = o2.nhc_new(*ens["temp"])
    ens["nvt_thermo"] = o2.run_nvt(x, v, np.zeros_like(x), atoms.type_ids, dt, steps, chain).tolist()
    ens["nvt_xi"] = list(chain.xi)
    o3 = Oracle.cubic(L)
    o3.insert(1, 1, 0.238, 3.405, rc)
    x, v = atoms.positions.copy(), atoms.velocities.copy()
    chain = o3.nhc_new(*ens["temp"])
    baro = o3.mtk_new(ens["iso"][0], ens["iso"][1], atoms.n_atoms, ens["temp"][0])
    th, htr = o3.run_npt(x, v, np.zeros_like(x), atoms.type_ids, dt, steps, baro, chain)
    ens["npt_thermo"], ens["npt_h"], ens["npt_momentum"] = th.tolist(), htr.tolist(), list(baro.momentum)
    ens["npt_positions_end"] = x.tolist()
    with open(os.path.join(HERE, "oracle_small_ensembles.json"), "w") as fh:
        json.dump(ens, fh)
    ref = "/root/reference/example/argon4000.txt"
    if os.path.exists(ref):
        with open(ref) as fh:
            lines = [ln.rstrip("\n") for ln in fh.readlines()[:32]]
        with open(os.path.join(HERE, "argon4000_head.json"), "w") as fh:
            json.dump({"source": "example/argon4000.txt (first 32 lines)", "lines": lines}, fh, indent=0)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
