"""A/B timing of kernel variants on the 4M-atom workload (run on the GPU box)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T0 = float(sys.argv[2]) if len(sys.argv) > 2 else 43.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
atoms = fcc_argon(ncell, temperature=T0, seed=12345)
for fv, bv, cd in ((3, 2, 1), (3, 3, 1), (3, 3, 2), (3, 2, 2)):
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.set_option("force_variant", fv)
    m.set_option("build_variant", bv)
    m.set_option("cell_div", cd)
    m.attach(atoms)
    m.compute()
    m.step_nve(0.25, 10)
    m.set_profiling(True)
    m.timings(reset=True)
    t0 = time.perf_counter()
    th = m.step_nve(0.25, steps)
    m.synchronize()
    dt = time.perf_counter() - t0
    tim = m.timings()
    st = m.stats()
    print(json.dumps({"force_variant": fv, "build_variant": bv, "cell_div": cd, "ms_per_step": 1e3 * dt / steps,
                      "atom_steps_per_s": atoms.n_atoms * steps / dt,
                      "per_launch_ms": {k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in tim.items() if v["launches"]},
                      "per_step_ms": {k: round(v["ms"] / steps, 4) for k, v in tim.items() if v["launches"]},
                      "builds": st["n_builds"], "pe_last": float(th["pe"][-1])}))
    m.close()
