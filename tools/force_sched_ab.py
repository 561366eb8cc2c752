"""A/B of two experiments on k_force_vv at 4M atoms: threads per block (force_block) and SM-aware work assignment
(force_sched).  One JSON line per combination; PE after the run must agree to the last bit (same lists, same per-thread sums;
the block partial sums change with the block size, so the PE may differ in the last digits between block sizes)."""
import json, os, sys, time
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T0 = float(sys.argv[2]) if len(sys.argv) > 2 else 43.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 48
combos = os.environ.get("FORCE_AB", "0,0;0,1;256,0;256,1;512,0;512,1;1024,0;1024,1")
for combo in combos.split(";"):
    fb, fs = (int(v) for v in combo.split(","))
    atoms = fcc_argon(ncell, temperature=T0, seed=12345)
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.set_option("force_block", fb)
    m.set_option("force_sched", fs)
    m.attach(atoms)
    m.compute()
    m.step_nve(0.25, 16)
    m.synchronize()
    t0 = time.perf_counter()
    m.step_nve(0.25, steps)
    m.synchronize()
    ms_graph = 1e3 * (time.perf_counter() - t0) / steps
    m.set_profiling(True)
    m.timings(reset=True)
    th = m.step_nve(0.25, steps)
    m.synchronize()
    tim = m.timings()
    m.set_profiling(False)
    print(json.dumps({"force_block": fb or 128, "force_sched": fs, "n_atoms": atoms.n_atoms, "ms_per_step_graphs": round(ms_graph, 4),
                      "force_ms_per_launch": round(tim["force"]["ms"] / max(tim["force"]["launches"], 1), 4),
                      "class_ms_per_step": {k: round(v["ms"] / steps, 4) for k, v in tim.items() if v["launches"]},
                      "pe_last": float(th["pe"][-1]).hex(), "ke_last": float(th["ke"][-1]).hex()}), flush=True)
    del m
