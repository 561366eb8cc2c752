"""ctypes loader for oracle/libpis_oracle.so -- the CPU ORACLE (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under pis_b200/ imports it.  PARITY UNPINNED by the reference (no rustc here,
reference tests pin no physics): see the header of pis_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "pis_oracle.c")
LIB = os.path.join(_HERE, "libpis_oracle.so")


class Box(C.Structure):
    _fields_ = [("h", C.c_double * 9), ("hinv", C.c_double * 9), ("pbc", C.c_int * 3)]


class Nhc(C.Structure):
    _fields_ = [("chain_size", C.c_int32), ("pad", C.c_int32), ("start_temperature", C.c_double),
                ("end_temperature", C.c_double), ("target_temperature", C.c_double), ("xi", C.c_double * 3),
                ("eta", C.c_double * 3), ("g", C.c_double * 3), ("q", C.c_double * 3)]


class Mtk(C.Structure):
    _fields_ = [("target_pressure", C.c_double * 9), ("momentum", C.c_double * 9), ("w", C.c_double)]


class Table(C.Structure):
    _fields_ = [("n_types", C.c_int), ("eps", C.c_void_p), ("sigma", C.c_void_p), ("rcut", C.c_void_p),
                ("present", C.c_void_p), ("shift", C.c_int)]


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC",
                        "-shared", "-fvisibility=hidden", "-o", LIB, SRC, "-lm"], check=True, cwd=_HERE)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        vp, d, i64 = C.c_void_p, C.c_double, C.c_int64
        bp, tp = C.POINTER(Box), C.POINTER(Table)
        sig = {
            "orc_box_new": (C.c_int, [vp, vp, bp]),
            "orc_box_from_lammps": (C.c_int, [d] * 9 + [bp]),
            "orc_box_volume": (d, [bp]),
            "orc_min_image": (None, [bp, vp]),
            "orc_wrap_pos": (None, [bp, vp]),
            "orc_max_rcut": (d, [tp]),
            "orc_divide_into_cells": (None, [bp, d, vp]),
            "orc_rcut_cells": (None, [bp, vp, i64, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp, vp]),
            "orc_forward_offsets": (vp, []),
            "orc_lj_pair": (None, [d, d, d, C.c_int, vp, vp, vp]),
            "orc_set_quiet": (None, [C.c_int]),
            "orc_compute_potential": (d, [bp, tp, i64, vp, vp, vp]),
            "orc_compute_potential_omp": (d, [bp, tp, i64, vp, vp, vp, C.c_int]),
            "orc_compute_potential_n2": (d, [bp, tp, i64, vp, vp, vp]),
            "orc_build_neighbour_list": (i64, [bp, tp, i64, vp, vp, d, vp, vp, i64]),
            "orc_build_neighbour_list_omp": (i64, [bp, tp, i64, vp, vp, d, vp, vp, i64, C.c_int]),
            "orc_compute_potential_list": (d, [bp, tp, i64, vp, vp, vp, vp, vp]),
            "orc_kinetic_energy": (d, [i64, vp, vp, vp]),
            "orc_temperature": (d, [i64, d]),
            "orc_virial_trace": (d, [i64, vp, vp]),
            "orc_pressure": (d, [bp, i64, vp, vp, d]),
            "orc_verlet_step_nve": (d, [bp, tp, i64, vp, vp, vp, vp, vp, d, C.c_int, C.c_int]),
            "orc_run_nve": (None, [bp, tp, i64, vp, vp, vp, vp, vp, d, i64, C.c_int, C.c_int, vp]),
            "orc_nhc_new": (None, [C.POINTER(Nhc), d, d, d]),
            "orc_nhc_kinetic_energy": (d, [C.POINTER(Nhc)]),
            "orc_verlet_step_nvt_nhc": (d, [bp, tp, i64, vp, vp, vp, vp, vp, d, C.POINTER(Nhc), C.c_int, C.c_int]),
            "orc_run_nvt": (None, [bp, tp, i64, vp, vp, vp, vp, vp, d, i64, C.POINTER(Nhc), C.c_int, C.c_int, vp]),
            "orc_mat3_exp": (None, [vp, vp]),
            "orc_pressure_tensor": (None, [bp, i64, vp, vp, vp, vp]),
            "orc_mtk_new": (None, [C.POINTER(Mtk), vp, d, i64, d]),
            "orc_mtk_scale": (None, [C.POINTER(Mtk), d, C.c_int, vp]),
            "orc_mtk_kinetic_energy": (d, [C.POINTER(Mtk)]),
            "orc_mtk_potential_energy": (d, [C.POINTER(Mtk), vp]),
            "orc_scale_box": (C.c_int, [bp, vp, i64, vp]),
            "orc_verlet_step_npt_mtk": (d, [bp, tp, i64, vp, vp, vp, vp, vp, d, C.POINTER(Mtk), C.POINTER(Nhc), C.c_int,
                                            C.c_int]),
            "orc_run_npt": (None, [bp, tp, i64, vp, vp, vp, vp, vp, d, i64, C.POINTER(Mtk), C.POINTER(Nhc), C.c_int,
                                   C.c_int, vp, vp]),
            "orc_rcut_threshold": (d, [d]),
            "orc_max_threads": (C.c_int, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Convenience wrapper: one box + one pair table + masses."""

    def __init__(self, h_colmajor, pbc=(1, 1, 1), n_types=1, masses=(39.948,), shift=True):
        self.lib = load()
        self.box = Box()
        h = np.ascontiguousarray(h_colmajor, dtype=np.float64).reshape(9)
        pb = np.array([int(p) for p in pbc], dtype=np.int32)
        if self.lib.orc_box_new(_p(h), _p(pb), C.byref(self.box)) != 0:
            raise ValueError("Box matrix should be invertible")
        self.n_types = n_types
        self.eps = np.zeros((n_types, n_types))
        self.sigma = np.zeros((n_types, n_types))
        self.rcut = np.zeros((n_types, n_types))
        self.present = np.zeros((n_types, n_types), dtype=np.uint8)
        self.masses = np.ascontiguousarray(masses, dtype=np.float64)
        self.table = Table(n_types, _p(self.eps).value, _p(self.sigma).value, _p(self.rcut).value,
                           _p(self.present).value, int(shift))
        self.lib.orc_set_quiet(1)

    @classmethod
    def cubic(cls, L, **kw):
        return cls([L, 0, 0, 0, L, 0, 0, 0, L], **kw)

    def insert(self, i, j, eps, sigma, rcut):
        self.eps[i - 1, j - 1], self.sigma[i - 1, j - 1], self.rcut[i - 1, j - 1] = eps, sigma, rcut
        self.present[i - 1, j - 1] = 1

    @property
    def h(self):
        return np.array(self.box.h[:]).reshape(3, 3).T

    @property
    def hinv(self):
        return np.array(self.box.hinv[:]).reshape(3, 3).T

    def max_rcut(self):
        return self.lib.orc_max_rcut(C.byref(self.table))

    def divide_into_cells(self, rcut):
        n = np.zeros(3, dtype=np.uint64)
        self.lib.orc_divide_into_cells(C.byref(self.box), float(rcut), _p(n))
        return tuple(int(x) for x in n)

    def rcut_cells(self, pos, nx, ny, nz):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = pos.shape[0]
        cell_of = np.zeros(n, dtype=np.int64)
        start = np.zeros(nx * ny * nz + 1, dtype=np.int64)
        atoms = np.zeros(n, dtype=np.int64)
        self.lib.orc_rcut_cells(C.byref(self.box), _p(pos), n, nx, ny, nz, _p(cell_of), _p(start), _p(atoms))
        return cell_of, start, atoms

    def min_image(self, d):
        d = np.array(d, dtype=np.float64)
        self.lib.orc_min_image(C.byref(self.box), _p(d))
        return d

    def wrap(self, r):
        r = np.array(r, dtype=np.float64)
        self.lib.orc_wrap_pos(C.byref(self.box), _p(r))
        return r

    def lj_pair(self, eps, sigma, rcut, rij, shift=True):
        rij = np.ascontiguousarray(rij, dtype=np.float64)
        u = C.c_double()
        f = np.zeros(3)
        self.lib.orc_lj_pair(eps, sigma, rcut, int(shift), _p(rij), C.byref(u), _p(f))
        return u.value, f

    def compute_potential(self, pos, types, forces=None, mode="serial", threads=0):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        types = np.ascontiguousarray(types, dtype=np.int32)
        n = pos.shape[0]
        f = np.zeros((n, 3)) if forces is None else forces
        b, t = C.byref(self.box), C.byref(self.table)
        if mode == "serial":
            pe = self.lib.orc_compute_potential(b, t, n, _p(pos), _p(types), _p(f))
        elif mode == "omp":
            pe = self.lib.orc_compute_potential_omp(b, t, n, _p(pos), _p(types), _p(f), threads)
        elif mode == "n2":
            pe = self.lib.orc_compute_potential_n2(b, t, n, _p(pos), _p(types), _p(f))
        else:
            raise ValueError(mode)
        return pe, f

    def build_neighbour_list(self, pos, types, extra=0.0, mode="serial", threads=0, cap_per_atom=64):
        """CSR (start, nbr) of the full list; mode="omp" distributes the cells over threads (identical rows)."""
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        types = np.ascontiguousarray(types, dtype=np.int32)
        n = pos.shape[0]
        start = np.zeros(n + 1, dtype=np.int64)
        cap = max(int(cap_per_atom) * n, 1024)
        while True:
            nbr = np.empty(cap, dtype=np.int32)
            if mode == "omp":
                tot = self.lib.orc_build_neighbour_list_omp(C.byref(self.box), C.byref(self.table), n, _p(pos), _p(types),
                                                            float(extra), _p(start), _p(nbr), cap, int(threads))
            else:
                tot = self.lib.orc_build_neighbour_list(C.byref(self.box), C.byref(self.table), n, _p(pos), _p(types),
                                                        float(extra), _p(start), _p(nbr), cap)
            if tot >= 0:
                return start, nbr[:tot]
            cap *= 4

    def compute_potential_list(self, pos, types, start, nbr):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        types = np.ascontiguousarray(types, dtype=np.int32)
        f = np.zeros_like(pos)
        pe = self.lib.orc_compute_potential_list(C.byref(self.box), C.byref(self.table), pos.shape[0], _p(pos),
                                                 _p(types), _p(start), _p(nbr), _p(f))
        return pe, f

    def kinetic_energy(self, vel, types):
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        types = np.ascontiguousarray(types, dtype=np.int32)
        return self.lib.orc_kinetic_energy(vel.shape[0], _p(vel), _p(types), _p(self.masses))

    def temperature(self, n, ke):
        return self.lib.orc_temperature(n, ke)

    def virial_trace(self, pos, forces):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        forces = np.ascontiguousarray(forces, dtype=np.float64)
        return self.lib.orc_virial_trace(pos.shape[0], _p(pos), _p(forces))

    def pressure(self, pos, forces, ke):
        return self.lib.orc_pressure(C.byref(self.box), pos.shape[0], _p(pos), _p(forces), ke)

    def verlet_step_nve(self, pos, vel, forces, types, dt, mode="serial", threads=0):
        """In place on pos/vel/forces ((N,3) float64 C-contiguous). Returns PE(t+dt)."""
        for a in (pos, vel, forces):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        types = np.ascontiguousarray(types, dtype=np.int32)
        return self.lib.orc_verlet_step_nve(C.byref(self.box), C.byref(self.table), pos.shape[0], _p(pos), _p(vel),
                                            _p(forces), _p(types), _p(self.masses), float(dt),
                                            0 if mode == "serial" else 1, threads)

    def run_nve(self, pos, vel, forces, types, dt, steps, mode="serial", threads=0):
        """Simulation::run NVE arm. Returns thermo[(steps+1), 5] = PE, KE, H, T, P (row 0: step-0 PE)."""
        types = np.ascontiguousarray(types, dtype=np.int32)
        thermo = np.zeros((steps + 1, 5))
        self.lib.orc_run_nve(C.byref(self.box), C.byref(self.table), pos.shape[0], _p(pos), _p(vel), _p(forces),
                             _p(types), _p(self.masses), float(dt), steps, 0 if mode == "serial" else 1, threads,
                             _p(thermo))
        return thermo

    def nhc_new(self, start_temperature, end_temperature, tau):
        c = Nhc()
        self.lib.orc_nhc_new(C.byref(c), float(start_temperature), float(end_temperature), float(tau))
        return c

    def run_nvt(self, pos, vel, forces, types, dt, steps, nhc, mode="serial", threads=0):
        """Simulation::run NVT arm. thermo[(steps+1), 5] = PE, KE, H (incl. thermostat energy), T, P."""
        types = np.ascontiguousarray(types, dtype=np.int32)
        thermo = np.zeros((steps + 1, 5))
        self.lib.orc_run_nvt(C.byref(self.box), C.byref(self.table), pos.shape[0], _p(pos), _p(vel), _p(forces),
                             _p(types), _p(self.masses), float(dt), steps, C.byref(nhc), 0 if mode == "serial" else 1,
                             threads, _p(thermo))
        return thermo

    def mat3_exp(self, a_colmajor):
        """nalgebra Matrix3::exp restated (column-major 9-vector in, 9-vector out)."""
        a = np.ascontiguousarray(a_colmajor, dtype=np.float64).reshape(9)
        out = np.zeros(9)
        self.lib.orc_mat3_exp(_p(a), _p(out))
        return out

    def pressure_tensor(self, pos, vel, forces):
        out = np.zeros(9)
        self.lib.orc_pressure_tensor(C.byref(self.box), pos.shape[0], _p(pos), _p(vel), _p(forces), _p(out))
        return out

    def box_volume(self):
        return self.lib.orc_box_volume(C.byref(self.box))

    def scale_box(self, scale, pos):
        """Atoms::scale_box (transformations.rs:6-15): box and positions in place; scale is a 3x3 (row, col) array."""
        sc = np.ascontiguousarray(np.asarray(scale, dtype=np.float64).T.reshape(9))
        if self.lib.orc_scale_box(C.byref(self.box), _p(sc), pos.shape[0], _p(pos)) != 0:
            raise ValueError("Box matrix should be invertible")

    def mtk_new(self, target_pressure, tau, n_atoms, target_temperature):
        """MTKBarostat::new_from_args (npt.rs:67-88): `iso p` gives target = p * identity."""
        m = Mtk()
        tp = np.ascontiguousarray(np.eye(3) * float(target_pressure) if np.isscalar(target_pressure) else target_pressure,
                                  dtype=np.float64).reshape(9)
        self.lib.orc_mtk_new(C.byref(m), _p(tp), float(tau), int(n_atoms), float(target_temperature))
        return m

    def run_npt(self, pos, vel, forces, types, dt, steps, mtk, nhc, mode="serial", threads=0):
        """Simulation::run NPT arm; the box of this Oracle is updated in place.
        Returns (thermo[(steps+1), 5], h_trace[(steps+1), 9])."""
        types = np.ascontiguousarray(types, dtype=np.int32)
        thermo = np.zeros((steps + 1, 5))
        htr = np.zeros((steps + 1, 9))
        self.lib.orc_run_npt(C.byref(self.box), C.byref(self.table), pos.shape[0], _p(pos), _p(vel), _p(forces),
                             _p(types), _p(self.masses), float(dt), steps, C.byref(mtk), C.byref(nhc),
                             0 if mode == "serial" else 1, threads, _p(thermo), _p(htr))
        return thermo, htr

    def rcut_threshold(self, rc):
        return self.lib.orc_rcut_threshold(float(rc))

    def max_threads(self):
        return self.lib.orc_max_threads()
