"""Does batching cost anything?  The same 60 NVE steps from the same state as one batch, as 6 batches of 10, and as 6 batches of
10 with an asynchronous position download per batch (the C++ host's Simulation::run loop)."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
from pis_b200 import Atoms, LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
base = fcc_argon(ncell, temperature=43.0, seed=12345, pinned=True)
m0 = LJCudaManager(skin=0.3 * 3.405)
m0.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
m0.attach(base); m0.compute(); m0.step_nve(0.25, 120); m0.download(base); m0.close()   # a thermalised state
x0, v0, f0 = base.positions.copy(), base.velocities.copy(), base.forces.copy()
for mode in ("one_batch", "batches_of_10", "batches_of_10_with_download", "one_batch"):
    base.positions[...], base.velocities[...], base.forces[...] = x0, v0, f0
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.attach(base); m.compute(); m.step_nve(0.25, 10); m.synchronize()
    if mode.endswith("download"):
        m.download_begin(base, positions=True, velocities=False, forces=False); m.download_end()
    b0 = m.stats()["n_builds"]
    per = []
    t0 = time.perf_counter()
    if mode == "one_batch":
        m.step_nve(0.25, 60)
    else:
        for _ in range(6):
            t1 = time.perf_counter()
            m.step_nve(0.25, 10)
            if mode.endswith("download"):
                m.download_end()
                m.download_begin(base, positions=True, velocities=False, forces=False)
            per.append(round(1e3 * (time.perf_counter() - t1), 3))
        if mode.endswith("download"):
            m.download_end()
    m.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    print(json.dumps({"mode": mode, "ms_per_step": round(ms / 60, 4), "builds": m.stats()["n_builds"] - b0, "per_batch_ms": per}), flush=True)
    m.close()
