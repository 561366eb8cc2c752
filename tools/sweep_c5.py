"""BASELINE config 5: cutoff / density sweep at ~4M atoms (rc in {2.5,3,4,5} sigma x rho* in {0.6,0.8,1.0}).
FCC lattice with a = sigma (4/rho*)^(1/3); reports K (list), K_in, ms/step per kernel class, atom-steps/s and the
force kernel's HBM / FP64 numbers.  Run on the GPU box; writes one JSON line per point."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon

SIG, EPS = 3.405, 0.238
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T0 = float(sys.argv[3]) if len(sys.argv) > 3 else 43.0
RHOS = tuple(float(v) for v in sys.argv[4].split(",")) if len(sys.argv) > 4 else (1.0, 0.8, 0.6)
for rho in RHOS:
    a = SIG * (4.0 / rho) ** (1.0 / 3.0)
    for rcs in (2.5, 3.0, 4.0, 5.0):
        rc, skin = rcs * SIG, 0.3 * SIG
        atoms = fcc_argon(ncell, temperature=T0, seed=12345, a=a)
        m = LJCudaManager(skin=skin)
        m.insert((1, 1), LennardJones(EPS, SIG, rc))
        try:
            m.attach(atoms); m.compute(); m.step_nve(0.25, 5)
            m.set_profiling(True); m.timings(reset=True)
            t0 = time.perf_counter(); th = m.step_nve(0.25, steps); m.synchronize(); dt = time.perf_counter() - t0
            tim = m.timings(); st = m.stats()
            nn = np.zeros(atoms.n_atoms, dtype=np.int32)
            from pis_b200 import capi
            capi.check(m._h, capi.load().pisb_neighbours(m._h, capi._ptr(nn), None, 0))
            K = float(nn.mean())
            f_ms = tim["force"]["ms"] / tim["force"]["launches"]
            out = {"rho*": rho, "rc/sigma": rcs, "n_atoms": atoms.n_atoms, "K_list": round(K, 1),
                   "ms_per_step": round(1e3 * dt / steps, 3), "atom_steps_per_s": atoms.n_atoms * steps / dt,
                   "force_ms": round(f_ms, 3), "force_GBps_algorithmic": round((48 + 4 * K) * atoms.n_atoms / f_ms / 1e6, 1),
                   "force_pairs_per_s": atoms.n_atoms * K / (f_ms * 1e-3),
                   "build_ms_per_step": round(tim["build"]["ms"] / steps, 3), "builds": st["n_builds"],
                   "list_capacity": st["list_capacity"], "device_GB": round(st["device_bytes"] / 1e9, 2),
                   "energy_drift_rel": float(np.abs((th["pe"] + th["ke"]) - (th["pe"][0] + th["ke"][0])).max() / abs(th["pe"][0] + th["ke"][0]))}
        except Exception as e:  # noqa: BLE001
            out = {"rho*": rho, "rc/sigma": rcs, "error": str(e)[:200]}
        print(json.dumps(out), flush=True)
        m.close()
