"""Diagnostic: the single-GPU half of the multi-GPU equivalence check's NVT leg (NVE steps, then a 60 K -> 90 K chain)."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.decomposition import create_velocities_distributed
from pis_b200.lattice import ARGON, fcc_argon

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fv = int(sys.argv[3]) if len(sys.argv) > 3 else 3
T0 = 60.0
n = 4 * ncell ** 3
ref = fcc_argon(ncell, temperature=0.0)
ref.velocities[...] = create_velocities_distributed(np.arange(n), np.full(n, ARGON["mass"]), T0, 777, n)
m = LJCudaManager(skin=0.3 * ARGON["sigma"])
m.insert((1, 1), LennardJones(ARGON["epsilon"], ARGON["sigma"], 2.5 * ARGON["sigma"], True))
if fv:
    m.set_option("force_variant", fv)
m.attach(ref)
m.compute()
th = m.step_nve(0.25, steps)
chain = m.nhc_new(T0, 1.5 * T0, 25.0)
th2, en = m.step_nvt_nhc(0.25, 30, chain, 0, 30)
m.download(ref)
print(json.dumps({"ncell": ncell, "steps": steps, "fv": fv, "nve_ke": th["ke"][[0, -1]].tolist(), "nvt_ke": th2["ke"].tolist()[:6] + th2["ke"].tolist()[-3:],
                  "nvt_pe": th2["pe"].tolist()[:3], "en": en.tolist()[:3], "vmax": float(np.abs(ref.velocities).max()),
                  "xi": list(chain.xi), "builds": m.stats()["n_builds"]}))
