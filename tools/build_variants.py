"""A/B of k_build_list_v3 experiment switches (option build_minb = MINB + 100 * pair records in flight + 1000 * row prefetch):
list equality with the default on a small periodic system, then ms per build at 4M atoms."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from pis_b200 import LennardJones, LJCudaManager, capi
from pis_b200.lattice import fcc_argon

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "8,10,1008,1010,210,1210,808,1808").split(",")]


def manager(minb):
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.set_option("build_minb", minb)
    return m


small = fcc_argon(12, temperature=60.0, seed=3, jitter=0.3)
ref_rows = None
for minb in variants:
    m = manager(minb)
    m.attach(small)
    rows = m.neighbours(small.n_atoms)
    if ref_rows is None:
        ref_rows = rows
    same = all(np.array_equal(a, b) for a, b in zip(rows, ref_rows))
    m.close()
    atoms = fcc_argon(ncell, temperature=43.0, seed=12345)
    m = manager(minb)
    m.attach(atoms)
    m.compute()
    m.step_nve(0.25, 60)      # melt a little: builds in the disordered state are the expensive ones
    m.set_profiling(True)
    m.timings(reset=True)
    b0 = m.stats()["n_builds"]
    th = m.step_nve(0.25, 60)
    tim = m.timings()
    nb = m.stats()["n_builds"] - b0
    m.set_profiling(False)
    nn = np.zeros(atoms.n_atoms, dtype=np.int32)
    capi.check(m._h, capi.load().pisb_neighbours(m._h, capi._ptr(nn), None, 0))
    print(json.dumps({"build_minb": minb, "small_lists_equal_default": bool(same), "builds": nb,
                      "ms_per_build": round(tim["build"]["ms"] / max(nb, 1), 4), "neighbours_total": int(nn.sum()),
                      "pe_last": float(th["pe"][-1])}), flush=True)
    m.close()
