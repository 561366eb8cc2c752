"""The 108 000-atom leg of smoke() without the oracle (compute-sanitizer target: the kernels of the 4M-atom benchmark --
k_build_list_v3, k_force_vv, the rebuild chain -- on a system small enough for racecheck).  argv[1]: cuda_graphs 0 / 1."""
import sys
sys.path.insert(0, ".")
import numpy as np
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import ARGON, fcc_argon

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ncell = int(sys.argv[2]) if len(sys.argv) > 2 else 30
atoms = fcc_argon(ncell, temperature=43.0, seed=12, jitter=0.05)
m = LJCudaManager(skin=0.3 * ARGON["sigma"], device=0)
m.insert((1, 1), LennardJones(ARGON["epsilon"], ARGON["sigma"], 2.5 * ARGON["sigma"], True))
m.set_option("cuda_graphs", graphs)
m.attach(atoms)
pe0 = m.compute()
th = m.step_nve(0.25, 12)
m.download(atoms)
print("sanitize_large: graphs", graphs, "n", atoms.n_atoms, "pe0", pe0, "pe_last", float(th["pe"][-1]), "builds", m.stats()["n_builds"],
      "finite", bool(np.isfinite(atoms.forces).all()))
m.close()
