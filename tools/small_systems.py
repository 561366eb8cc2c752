"""Throughput across system sizes (BASELINE config 0: 4000 atoms ... config 1: 256k atoms) for the step-kernel choices:
force_variant 0 = the automatic choice (8 / 4 lanes per atom + k_vv below 32k / 75k atoms, fused thread-per-atom above),
3 = fused thread-per-atom at any size, 7 = fused four-lanes-per-atom (k_force_q) at any size.  CUDA-graph replays."""
import json, sys, time
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon
sizes = ((10, 8.5, 2048), (20, 2.5 * 3.405, 1024), (30, 2.5 * 3.405, 512), (40, 2.5 * 3.405, 512))
for ncell, rc, steps in sizes:
    for T0 in (5.0, 43.0):
        for fv in (0, 3, 7):
            atoms = fcc_argon(ncell, temperature=T0, seed=12345)
            m = LJCudaManager(skin=0.3 * 3.405)
            m.set_option("force_variant", fv)
            m.insert((1, 1), LennardJones(0.238, 3.405, rc))
            m.attach(atoms); m.compute(); m.step_nve(0.25, 64)
            l0 = m.stats()["n_launches"]
            t0 = time.perf_counter(); m.step_nve(0.25, steps); m.synchronize(); dt = time.perf_counter() - t0
            st = m.stats()
            print(json.dumps({"n_atoms": atoms.n_atoms, "T0": T0, "force_variant": fv, "steps": steps,
                              "us_per_step": round(1e6 * dt / steps, 1), "atom_steps_per_s": atoms.n_atoms * steps / dt,
                              "builds": st["n_builds"], "kernels_per_step": round((st["n_launches"] - l0) / steps, 2)}), flush=True)
            m.close()
