"""A/B of the fused step kernel (option fuse_vv = 0: k_force_v3 + k_vv, 1: k_force_vv), graph replay + per-kernel pass; one
JSON line per variant.  (profiles/r01_fused_step.jsonl also holds the cache-hint variants 2/4/6/8 that were tried and
removed.)"""
import json, sys, time
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T0 = float(sys.argv[2]) if len(sys.argv) > 2 else 43.0
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
steps = 100
for fv in variants:
    atoms = fcc_argon(ncell, temperature=T0, seed=12345)
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.set_option("fuse_vv", fv)
    m.attach(atoms)
    m.compute()
    m.step_nve(0.25, 10)
    m.synchronize()
    t0 = time.perf_counter()
    m.step_nve(0.25, steps)
    m.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    m.set_profiling(True)
    m.timings(reset=True)
    m.step_nve(0.25, steps)
    tim = m.timings()
    m.set_profiling(False)
    per = {k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in tim.items() if v["launches"]}
    print(json.dumps({"fuse_vv": fv, "n_atoms": atoms.n_atoms, "T0": T0, "ms_per_step_graphs": round(ms, 4),
                      "ms_per_launch": per, "builds": m.stats()["n_builds"]}), flush=True)
    m.close()
