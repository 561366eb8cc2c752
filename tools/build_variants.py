"""A/B of k_build_list_v3 compiled for different occupancies (option build_minb), 4M atoms: ms per build."""
import json, sys
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for minb in [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "8,6,10,12").split(",")]:
    atoms = fcc_argon(ncell, temperature=43.0, seed=12345)
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.set_option("build_minb", minb)
    m.attach(atoms)
    m.compute()
    m.step_nve(0.25, 60)      # melt a little: builds in the disordered state are the expensive ones
    m.set_profiling(True)
    m.timings(reset=True)
    b0 = m.stats()["n_builds"]
    m.step_nve(0.25, 60)
    tim = m.timings()
    nb = m.stats()["n_builds"] - b0
    m.set_profiling(False)
    print(json.dumps({"build_minb": minb, "builds": nb, "ms_per_build": round(tim["build"]["ms"] / max(nb, 1), 4),
                      "ms_per_step_force": round(tim["force"]["ms"] / 60, 4)}), flush=True)
    m.close()
