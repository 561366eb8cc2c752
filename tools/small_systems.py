"""Throughput at small N (BASELINE config 0: 4000 atoms, config 1: 256k atoms), CUDA-graph replay vs classic launches."""
import json, sys, time
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon
for ncell, rc, steps in ((10, 8.5, 2048), (40, 2.5 * 3.405, 512), (100, 2.5 * 3.405, 128)):
    for T0 in (5.0, 43.0):
        for graphs in (1, 0):
            atoms = fcc_argon(ncell, temperature=T0, seed=12345)
            m = LJCudaManager(skin=0.3 * 3.405)
            m.set_option("cuda_graphs", graphs)
            m.insert((1, 1), LennardJones(0.238, 3.405, rc))
            m.attach(atoms); m.compute(); m.step_nve(0.25, 64)
            l0 = m.stats()["n_launches"]
            t0 = time.perf_counter(); m.step_nve(0.25, steps); m.synchronize(); dt = time.perf_counter() - t0
            st = m.stats()
            print(json.dumps({"n_atoms": atoms.n_atoms, "T0": T0, "cuda_graphs": graphs, "steps": steps,
                              "us_per_step": round(1e6 * dt / steps, 1), "atom_steps_per_s": atoms.n_atoms * steps / dt,
                              "builds": st["n_builds"], "kernels_per_step": round((st["n_launches"] - l0) / steps, 2)}), flush=True)
            m.close()
