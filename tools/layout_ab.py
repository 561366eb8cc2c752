"""A/B of the list-layout options on the 4M-atom workload: build_window (per-row x window in the v3 build) and list_align
(rows of a warp padded to a common length per stencil plane / row).  One JSON line per combination: ms per step (graph
replay), per-launch times of the kernel classes (profiled pass), list statistics, and the PE after the run (must agree to
the last bit between layouts: same neighbours, same entry order).
usage: python tools/layout_ab.py [ncell=100] [T0=43] [steps=48]   (combinations from LAYOUT_AB="align,window;..." or the default set)"""
import json, os, sys, time
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T0 = float(sys.argv[2]) if len(sys.argv) > 2 else 43.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 48
combos = os.environ.get("LAYOUT_AB", "0,0;0,1;1,1;2,1")
base = fcc_argon(ncell, temperature=T0, seed=12345)
for combo in combos.split(";"):
    align, window = (int(v) for v in combo.split(","))
    atoms = fcc_argon(ncell, temperature=T0, seed=12345) if combo != combos.split(";")[0] else base
    m = LJCudaManager(skin=0.3 * 3.405)
    m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
    m.set_option("list_align", align)
    m.set_option("build_window", window)
    m.attach(atoms)
    m.compute()
    m.step_nve(0.25, 16)          # thermalise: the first rebuilds, graph capture
    m.synchronize()
    t0 = time.perf_counter()
    m.step_nve(0.25, steps)
    m.synchronize()
    ms_graph = 1e3 * (time.perf_counter() - t0) / steps
    ls = m.list_stats()
    m.set_profiling(True)
    m.timings(reset=True)
    b0 = m.stats()["n_builds"]
    th = m.step_nve(0.25, steps)
    m.synchronize()
    tim = m.timings()
    st = m.stats()
    m.set_profiling(False)
    n = atoms.n_atoms
    print(json.dumps({"list_align": align, "build_window": window, "n_atoms": n, "ms_per_step_graphs": round(ms_graph, 4),
                      "per_launch_ms": {k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in tim.items() if v["launches"]},
                      "class_ms_per_step": {k: round(v["ms"] / steps, 4) for k, v in tim.items() if v["launches"]},
                      "builds_in_profiled_pass": st["n_builds"] - b0, "listed_per_atom": round(ls["listed"] / n, 2),
                      "in_range_per_atom": round(ls["in_range"] / n, 2), "index_words_per_atom": round(ls["index_words"] / n, 2),
                      "list_capacity": st["list_capacity"], "max_row": st["max_neighbours"], "device_GB": round(st["device_bytes"] / 1e9, 2),
                      "pe_last": float(th["pe"][-1]).hex()}), flush=True)
    del m
