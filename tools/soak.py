"""Long-run soak: many rebuilds, energy bookkeeping, momentum conservation (run on the GPU box)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon
for ncell, steps, T0 in ((40, 6000, 60.0), (100, 1500, 43.0)):
    atoms = fcc_argon(ncell, temperature=T0, seed=99)
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.attach(atoms); pe0 = m.compute()
    t0 = time.perf_counter()
    th = np.concatenate([m.step_nve(0.25, 500) for _ in range(steps // 500)])
    dt = time.perf_counter() - t0
    m.download(atoms)
    h = th["pe"] + th["ke"]
    p = (atoms.velocities * 39.948).sum(axis=0)
    out = {"n_atoms": atoms.n_atoms, "steps": steps, "T0": T0, "builds": m.stats()["n_builds"], "max_neighbours": m.stats()["max_neighbours"],
           "H_first": float(h[0]), "H_last": float(h[-1]), "H_drift_rel": float(np.abs(h - h[0]).max() / abs(h[0])),
           "T_last": float(atoms.temerature(th["ke"][-1])), "momentum_max": float(np.abs(p).max()),
           "finite": bool(np.isfinite(atoms.positions).all() and np.isfinite(h).all()),
           "in_box": bool((atoms.positions >= 0).all() and (atoms.positions <= atoms.sim_box.h[0, 0]).all()),
           "atom_steps_per_s": atoms.n_atoms * steps / dt}
    print(json.dumps(out), flush=True)
    m.close()
