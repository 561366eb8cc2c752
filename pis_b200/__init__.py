"""pis_b200 -- B200-native (sm_100a) implementation of the PIS per-timestep MD hot path
(cell binning -> Verlet neighbour list -> Lennard-Jones force/energy/virial -> velocity-Verlet NVE)
behind the reference's own operator interface.  See DESIGN.md / INTEGRATION.md."""
from .atoms import KB_KJPERMOLEKELVIN, Atoms
from .potentials import LennardJones, LJCudaManager
from .simulation_box import SimulationBox

__all__ = ["Atoms", "SimulationBox", "LennardJones", "LJCudaManager", "KB_KJPERMOLEKELVIN"]
