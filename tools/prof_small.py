"""A few classic-launch steps of the 4000-atom system (for an ncu launch list of the small-system regime)."""
import sys
sys.path.insert(0, ".")
from pis_b200 import LennardJones, LJCudaManager
from pis_b200.lattice import fcc_argon
atoms = fcc_argon(10, temperature=43.0, seed=12345)
m = LJCudaManager(skin=0.3 * 3.405)
m.set_option("cuda_graphs", 0)
m.insert((1, 1), LennardJones(0.238, 3.405, 8.5))
m.attach(atoms)
m.compute()
m.step_nve(0.25, 20)
print(m.stats())
