"""Run a few steps of one kernel-variant combination (for ncu captures)."""
import sys
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon

fv, bv = int(sys.argv[1]), int(sys.argv[2])
ncell = int(sys.argv[3]) if len(sys.argv) > 3 else 100
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
T0 = float(sys.argv[5]) if len(sys.argv) > 5 else 43.0
atoms = fcc_argon(ncell, temperature=T0, seed=12345)
m = LJCudaManager(skin=0.3 * 3.405)
m.insert((1, 1), LennardJones(0.238, 3.405, 2.5 * 3.405))
m.set_option("force_variant", fv)
m.set_option("build_variant", bv)
m.set_option("cell_div", int(sys.argv[6]) if len(sys.argv) > 6 else 0)
for kv in sys.argv[7:]:   # further library options name=value, e.g. cuda_graphs=0 (ncu cannot profile kernel nodes of graphs
    name, _, val = kv.partition("=")   # that hold conditional nodes), fuse_vv=0
    m.set_option(name, float(val))
m.attach(atoms)
m.compute()
m.step_nve(0.25, steps)
for _ in range(3):   # three full rebuild chains on the disordered state: launches steps+1 .. steps+3 of the build kernel are real builds
    m.invalidate_list()
    m.compute()
print(m.stats())
