"""Time one (force_variant, build_variant, cell_div) combination on the 4M-atom workload."""
import json, sys, time
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon
fv, bv, cd = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ncell = int(sys.argv[4]) if len(sys.argv) > 4 else 100
T0 = float(sys.argv[5]) if len(sys.argv) > 5 else 43.0
steps = int(sys.argv[6]) if len(sys.argv) > 6 else 40
atoms = fcc_argon(ncell, temperature=T0, seed=12345)
m = LJCudaManager(skin=0.3 * 3.405)
m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
m.set_option("force_variant", fv); m.set_option("build_variant", bv); m.set_option("cell_div", cd)
m.attach(atoms); m.compute(); m.step_nve(0.25, 10)
m.set_profiling(True); m.timings(reset=True)
t0 = time.perf_counter(); th = m.step_nve(0.25, steps); m.synchronize(); dt = time.perf_counter() - t0
tim = m.timings(); st = m.stats()
print(json.dumps({"fv": fv, "bv": bv, "cd": cd, "ms_per_step": round(1e3 * dt / steps, 4),
                  "per_launch_ms": {k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in tim.items() if v["launches"]},
                  "builds": st["n_builds"], "pe_last": float(th["pe"][-1])}))
