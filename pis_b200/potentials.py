"""Potential layer -- host mirror of src/potentials/{potential,lennard_jones,kind}.rs for the hot
path, backed by the sm_100a kernels through the C ABI.

`LJCudaManager` is the new `PotentialManagerKind` variant: it keeps the reference's table surface
(`insert`, `get`, `is_empty`, `max_rcut`, `get_potential_ij`) and the two trait methods on the path,
`compute_potential(atoms) -> f64` and `verlet_step_nve(atoms, dt) -> f64`, with the same argument
meaning (forces are ADDED by compute_potential; verlet_step_nve returns PE at the new positions).
There is no CPU implementation behind it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .atoms import Atoms


class LennardJones:
    """ref: src/potentials/lennard_jones.rs:14-30 (parameters only; the pair math runs on the GPU)."""

    def __init__(self, epsilon: float, sigma: float, rcut: float, shift: bool = True):
        self.epsilon, self.sigma, self.rcut, self.shift = float(epsilon), float(sigma), float(rcut), bool(shift)

    def get_rcut(self) -> float:
        return self.rcut


class LJCudaManager:
    """Device-backed LJ manager (drop-in for LJVOffsetManager, src/potentials/lennard_jones.rs:182-246)."""

    def __init__(self, skin: float = 0.0, device: int = 0, table: dict | None = None):
        self.table: dict[tuple[int, int], LennardJones] = dict(table or {})
        self.skin = float(skin)
        self.device = int(device)
        self._h = C.c_void_p(None)
        self._sig = None      # signature of (table, masses) the handle was created with
        self._box_sig = None
        self._atoms_id = None
        self._options: dict[str, float] = {}

    # ---- PairPotentialManager surface (src/potentials/potential.rs:148-193) -------------------
    @classmethod
    def new(cls, **kw):
        return cls(**kw)

    def is_empty(self) -> bool:
        return not self.table

    def insert(self, key, potential: LennardJones):
        self.table[(int(key[0]), int(key[1]))] = potential  # key stored AS GIVEN (system.rs:162)
        self._sig = None

    def get(self, key):
        return self.table.get((int(key[0]), int(key[1])))

    def max_rcut(self) -> float:
        m = 0.0
        for p in self.table.values():
            if m < p.get_rcut():
                m = p.get_rcut()
        return m

    def get_potential_ij(self, atoms: Atoms, i: int, j: int):
        ti, tj = int(atoms.type_ids[i]), int(atoms.type_ids[j])
        return self.get((ti, tj) if ti < tj else (tj, ti))

    # ---- handle management ---------------------------------------------------------------------
    def _ensure_handle(self, atoms: Atoms):
        lib = capi.load()
        nt = max(len(atoms.masses), max((max(k) for k in self.table), default=1))
        sig = (nt, tuple(atoms.masses), self.skin,
               tuple(sorted((k, p.epsilon, p.sigma, p.rcut, p.shift) for k, p in self.table.items())))
        if self._h and sig == self._sig:
            return
        self.close()
        mass = np.zeros(nt)
        mass[: len(atoms.masses)] = atoms.masses
        eps, sig_, rc = np.zeros((nt, nt)), np.zeros((nt, nt)), np.zeros((nt, nt))
        present = np.zeros((nt, nt), dtype=np.uint8)
        shift = True
        for (i, j), p in self.table.items():
            eps[i - 1, j - 1], sig_[i - 1, j - 1], rc[i - 1, j - 1] = p.epsilon, p.sigma, p.rcut
            present[i - 1, j - 1] = 1
            shift = p.shift
        h = C.c_void_p()
        rc_ = lib.pisb_create(self.device, nt, capi._ptr(mass), capi._ptr(eps), capi._ptr(sig_), capi._ptr(rc),
                              capi._ptr(present), int(shift), self.skin, C.byref(h))
        capi.check(None, rc_)
        self._h, self._sig, self._box_sig, self._atoms_id = h, sig, None, None
        for k, v in self._options.items():
            capi.check(self._h, lib.pisb_set_option(self._h, k.encode(), float(v)))

    def _ensure_box(self, atoms: Atoms):
        b = atoms.sim_box
        bsig = (b.h.tobytes(), b.h_inv.tobytes(), tuple(b.pbc))
        if bsig == self._box_sig:
            return
        pbc = np.array([1 if p else 0 for p in b.pbc], dtype=np.int32)
        hc, hic = b.h_colmajor(), b.h_inv_colmajor()
        capi.check(self._h, capi.load().pisb_set_box(self._h, capi._ptr(hc), capi._ptr(hic), capi._ptr(pbc)))
        self._box_sig = bsig

    def close(self):
        if self._h:
            capi.load().pisb_destroy(self._h)
            self._h = C.c_void_p(None)
            self._sig = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- PotentialManager trait, host-buffer (strict drop-in) form ----------------------------
    def compute_potential(self, atoms: Atoms) -> float:
        """ref: PotentialManager::compute_potential (src/potentials/potential.rs:13).
        Adds the LJ forces into atoms.forces and returns the total shifted PE."""
        self.attach(atoms, with_forces=True)
        pe = C.c_double()
        capi.check(self._h, capi.load().pisb_compute(self._h, 1, C.byref(pe)))
        self.download(atoms, positions=False, velocities=False, forces=True)
        return pe.value

    def verlet_step_nve(self, atoms: Atoms, dt: float) -> float:
        """ref: PotentialManager::verlet_step_nve (src/potentials/potential.rs:15-33), one step with
        HOST buffers: upload x, v, F; step on the device; download x, v, F; return PE(t+dt)."""
        self._ensure_handle(atoms)
        self._ensure_box(atoms)
        lib = capi.load()
        pe = C.c_double()
        fresh = self._atoms_id != (id(atoms), atoms.n_atoms)
        rc = lib.pisb_verlet_step_nve_host(self._h, atoms.n_atoms, capi._ptr(atoms.positions),
                                           capi._ptr(atoms.velocities), capi._ptr(atoms.forces),
                                           capi._ptr(atoms.type_ids) if fresh else None, float(dt), C.byref(pe))
        capi.check(self._h, rc)
        self._atoms_id = (id(atoms), atoms.n_atoms)
        return pe.value

    # ---- device-resident form (what Simulation::run uses) -------------------------------------
    def attach(self, atoms: Atoms, with_forces: bool = True):
        """Upload atoms (positions, velocities, forces, types) and the box."""
        self._ensure_handle(atoms)
        self._ensure_box(atoms)
        rc = capi.load().pisb_upload(self._h, atoms.n_atoms, capi._ptr(atoms.positions), capi._ptr(atoms.velocities),
                                     capi._ptr(atoms.forces) if with_forces else None, capi._ptr(atoms.type_ids))
        capi.check(self._h, rc)
        self._atoms_id = (id(atoms), atoms.n_atoms)

    def compute(self, accumulate: bool = False) -> float:
        pe = C.c_double()
        capi.check(self._h, capi.load().pisb_compute(self._h, int(accumulate), C.byref(pe)))
        return pe.value

    def step_nve(self, dt: float, nsteps: int = 1) -> np.ndarray:
        """nsteps x verlet_step_nve on the device; returns a structured array (pe, ke, virial_ref, virial_pair)."""
        out = np.zeros(int(nsteps), dtype=capi.THERMO_DTYPE)
        capi.check(self._h, capi.load().pisb_step_nve(self._h, float(dt), int(nsteps), capi._ptr(out)))
        return out

    @staticmethod
    def nhc_new(start_temperature: float, end_temperature: float, tau: float) -> "capi.Nhc":
        """NHThermostatChain::new_from_args (src/ensemble/nvt.rs:99-112): 3 links, target = start."""
        c = capi.Nhc()
        capi.check(None, capi.load().pisb_nhc_init(C.byref(c), float(start_temperature), float(end_temperature), float(tau)))
        return c

    def step_nvt_nhc(self, dt: float, nsteps: int, chain: "capi.Nhc", first_step: int, total_steps: int):
        """nsteps x (verlet_step_nvt_nhc + calculate_target_temperature) on the device (potential.rs:35-58,
        simulation.rs:53-57).  Returns (thermo records, thermostat energy per step); `chain` is updated in place."""
        out = np.zeros(int(nsteps), dtype=capi.THERMO_DTYPE)
        en = np.zeros(int(nsteps))
        capi.check(self._h, capi.load().pisb_step_nvt_nhc(self._h, float(dt), int(nsteps), C.byref(chain), int(first_step),
                                                          int(total_steps), capi._ptr(out), capi._ptr(en)))
        return out, en

    @staticmethod
    def mtk_new(target_pressure, tau: float, n_atoms: int, target_temperature: float) -> "capi.Mtk":
        """MTKBarostat::new_from_args (src/ensemble/npt.rs:24-43,67-88).  A scalar is the `iso p` form
        (target = p * identity, commands.rs:439); a 3x3 array is taken as is."""
        tp = np.eye(3) * float(target_pressure) if np.isscalar(target_pressure) else np.asarray(target_pressure, dtype=np.float64)
        tp = np.ascontiguousarray(tp.T.reshape(9))  # column-major
        m = capi.Mtk()
        capi.check(None, capi.load().pisb_mtk_init(C.byref(m), capi._ptr(tp), float(tau), int(n_atoms), float(target_temperature)))
        return m

    def step_npt_mtk(self, dt: float, nsteps: int, baro: "capi.Mtk", chain: "capi.Nhc", first_step: int, total_steps: int,
                     atoms: Atoms | None = None):
        """nsteps x (verlet_step_npt_mtk + calculate_target_temperature) (potential.rs:112-135, simulation.rs:58-62).
        Returns (thermo records, extended-system energy per step, box h per step as (nsteps, 3, 3)); `baro` and `chain`
        are updated in place, and so is `atoms.sim_box` when atoms is given (scale_box changes the box)."""
        out = np.zeros(int(nsteps), dtype=capi.THERMO_DTYPE)
        en = np.zeros(int(nsteps))
        h9 = np.zeros((int(nsteps), 9))
        capi.check(self._h, capi.load().pisb_step_npt_mtk(self._h, float(dt), int(nsteps), C.byref(baro), C.byref(chain),
                                                          int(first_step), int(total_steps), capi._ptr(out), capi._ptr(en),
                                                          capi._ptr(h9)))
        if atoms is not None:
            self.sync_box(atoms)
        return out, en, h9.reshape(-1, 3, 3).transpose(0, 2, 1)

    def get_box(self):
        """(h, h_inv) held by the device handle, as 3x3 arrays (row, column)."""
        hc, hic = np.zeros(9), np.zeros(9)
        capi.check(self._h, capi.load().pisb_get_box(self._h, capi._ptr(hc), capi._ptr(hic)))
        return hc.reshape(3, 3).T.copy(), hic.reshape(3, 3).T.copy()

    def sync_box(self, atoms: Atoms):
        """atoms.sim_box <- the handle's box (after NPT steps)."""
        h, hi = self.get_box()
        atoms.sim_box.h[...] = h
        atoms.sim_box.h_inv[...] = hi
        b = atoms.sim_box
        self._box_sig = (b.h.tobytes(), b.h_inv.tobytes(), tuple(b.pbc))

    def download(self, atoms: Atoms, positions=True, velocities=True, forces=True):
        rc = capi.load().pisb_download(self._h, capi._ptr(atoms.positions) if positions else None,
                                       capi._ptr(atoms.velocities) if velocities else None,
                                       capi._ptr(atoms.forces) if forces else None)
        capi.check(self._h, rc)

    def download_begin(self, atoms: Atoms, positions=True, velocities=False, forces=False):
        """Asynchronous download for dump steps (Simulation::run, src/simulation.rs:31-37): the device takes a snapshot in
        stream order and copies it on a second stream while the next batch runs; valid after download_end()."""
        rc = capi.load().pisb_download_begin(self._h, capi._ptr(atoms.positions) if positions else None,
                                             capi._ptr(atoms.velocities) if velocities else None,
                                             capi._ptr(atoms.forces) if forces else None)
        capi.check(self._h, rc)

    def download_end(self):
        capi.check(self._h, capi.load().pisb_download_end(self._h))

    def thermo_now(self) -> dict:
        t = capi.Thermo()
        capi.check(self._h, capi.load().pisb_thermo_now(self._h, C.byref(t)))
        return {"pe": t.pe, "ke": t.ke, "virial_ref": t.virial_ref, "virial_pair": t.virial_pair}

    def start_velocities(self, temperature: float, seed: int = 0):
        """Atoms::start_velocities (velocities.rs:10-15) on the device, for the attached atoms: Gaussian velocities
        (this repository's id-keyed generator), drift removal, rescale to `temperature`."""
        capi.check(self._h, capi.load().pisb_start_velocities(self._h, float(temperature), int(seed)))

    def neighbours(self, n_atoms: int):
        """Current Verlet list in original ids: list of sorted numpy arrays (test hook)."""
        lib = capi.load()
        nn = np.zeros(n_atoms, dtype=np.int32)
        capi.check(self._h, lib.pisb_neighbours(self._h, capi._ptr(nn), None, 0))
        cap = int(nn.max()) if n_atoms else 0
        nb = np.zeros((n_atoms, max(cap, 1)), dtype=np.int32)
        capi.check(self._h, lib.pisb_neighbours(self._h, capi._ptr(nn), capi._ptr(nb), max(cap, 1)))
        return [np.sort(nb[i, : nn[i]]) for i in range(n_atoms)]

    def neighbours_padded(self, n_atoms: int):
        """Current Verlet list in original ids as (counts, rows): rows[i, :counts[i]] are the neighbours of atom i in list
        order, the rest of the row is undefined (test hook for sizes where a Python list of 4M arrays is too slow)."""
        lib = capi.load()
        nn = np.zeros(n_atoms, dtype=np.int32)
        capi.check(self._h, lib.pisb_neighbours(self._h, capi._ptr(nn), None, 0))
        cap = max(int(nn.max()) if n_atoms else 0, 1)
        nb = np.empty((n_atoms, cap), dtype=np.int32)
        capi.check(self._h, lib.pisb_neighbours(self._h, capi._ptr(nn), capi._ptr(nb), cap))
        return nn, nb

    def list_stats(self) -> dict:
        """Listed pairs, pairs inside the cutoff now, index words stored (pisb_list_stats; device-side count)."""
        out = np.zeros(3, dtype=np.int64)
        capi.check(self._h, capi.load().pisb_list_stats(self._h, capi._ptr(out)))
        return {"listed": int(out[0]), "in_range": int(out[1]), "index_words": int(out[2])}

    def invalidate_list(self):
        capi.check(self._h, capi.load().pisb_invalidate_list(self._h))

    def stats(self) -> dict:
        s = capi.Stats()
        capi.check(self._h, capi.load().pisb_stats(self._h, C.byref(s)))
        return s.as_dict()

    def set_profiling(self, on: bool):
        capi.check(self._h, capi.load().pisb_set_profiling(self._h, int(on)))

    def timings(self, reset: bool = False) -> dict:
        ms = np.zeros(capi.K_COUNT)
        cnt = np.zeros(capi.K_COUNT, dtype=np.int64)
        capi.check(self._h, capi.load().pisb_timings(self._h, capi._ptr(ms), capi._ptr(cnt)))
        if reset:
            capi.check(self._h, capi.load().pisb_timings_reset(self._h))
        return {k: {"ms": float(ms[i]), "launches": int(cnt[i])} for i, k in enumerate(capi.K_NAMES)}

    def set_option(self, name: str, value: float):
        """Tuning knob (list_capacity, force_variant, build_variant); kept across handle re-creation."""
        self._options[name] = float(value)
        if self._h:
            capi.check(self._h, capi.load().pisb_set_option(self._h, name.encode(), float(value)))

    def synchronize(self):
        capi.check(self._h, capi.load().pisb_synchronize(self._h))

    @property
    def stream_ptr(self) -> int:
        return int(capi.load().pisb_stream(self._h) or 0)
