"""Where the host-buffer step (e2e) and the dump loop (e2e_resident) spend their time: raw PCIe rates, chunk-size sweep of
the pipelined step, per-batch times of the resident loop.  JSON lines."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n = 4 * ncell ** 3
# ---- raw PCIe: 3 x n doubles, pinned ----
hb = torch.empty(3 * n, dtype=torch.float64).pin_memory()
hb2 = torch.empty(3 * n, dtype=torch.float64).pin_memory()
db = torch.empty(3 * n, dtype=torch.float64, device="cuda")
db2 = torch.empty(3 * n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    db.copy_(hb, non_blocking=True)
def d2h():
    hb2.copy_(db2, non_blocking=True)
def both():
    with torch.cuda.stream(s1):
        db.copy_(hb, non_blocking=True)
    with torch.cuda.stream(s2):
        hb2.copy_(db2, non_blocking=True)
h2d(); d2h(); both()
gb = 24 * n / 1e9
print(json.dumps({"pcie": {"bytes": 24 * n, "h2d_GBps": gb / timed(h2d), "d2h_GBps": gb / timed(d2h),
                           "duplex_each_GBps": gb / timed(both)}}), flush=True)
del hb, hb2, db, db2

atoms = fcc_argon(ncell, temperature=43.0, seed=12345, pinned=True)
for pipe, chunk, fuse in ((0, 0, 1), (1, 0, 1), (1, n // 4, 1), (1, n // 16, 1), (1, n // 32, 1), (1, 0, 0)):
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.set_option("host_pipeline", pipe)
    m.set_option("host_chunk_atoms", chunk)
    m.set_option("fuse_vv", fuse)
    m.compute_potential(atoms)
    m.verlet_step_nve(atoms, 0.25)
    m.synchronize()
    per = []
    for _ in range(12):
        t0 = time.perf_counter()
        m.verlet_step_nve(atoms, 0.25)
        per.append(1e3 * (time.perf_counter() - t0))
    print(json.dumps({"host_step": {"pipeline": pipe, "chunk_atoms": chunk, "fuse_vv": fuse, "ms_median": float(np.median(per)),
                                    "ms_min": min(per), "ms_max": max(per)}}), flush=True)
    m.close()

for fuse in (1, 0):
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.set_option("fuse_vv", fuse)
    m.attach(atoms)
    m.compute()
    m.step_nve(0.25, 10)
    m.synchronize()
    rows = []
    for b in range(6):
        t0 = time.perf_counter()
        m.step_nve(0.25, 10)
        t1 = time.perf_counter()
        m.download_end()
        t2 = time.perf_counter()
        m.download_begin(atoms, positions=True)
        t3 = time.perf_counter()
        rows.append([round(1e3 * (t1 - t0), 3), round(1e3 * (t2 - t1), 3), round(1e3 * (t3 - t2), 3)])
    m.download_end()
    st = m.stats()
    print(json.dumps({"resident_loop": {"fuse_vv": fuse, "per_batch_ms[step_nve(10), download_end, download_begin]": rows,
                                        "n_builds": st["n_builds"], "n_launches": st["n_launches"]}}), flush=True)
    m.close()
