"""ctypes binding of libpisb200.so (include/pisb200.h).  Plumbing only: every compute call goes
to the hand-written sm_100a kernels; there is no Python/CPU fallback -- a missing library or a
missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PISB_LIB") or os.path.join(_HERE, "libpisb200.so")  # PISB_LIB: A/B builds (tools/)

PISB_OK, PISB_ERR_INVALID, PISB_ERR_CUDA, PISB_ERR_NO_DEVICE, PISB_ERR_CAPACITY, PISB_ERR_STATE, PISB_ERR_COMM = range(7)
K_NAMES = ["integrate", "bin", "sort", "build", "force", "reduce", "halo", "copy"]
K_COUNT = len(K_NAMES)


class Thermo(C.Structure):
    _fields_ = [("pe", C.c_double), ("ke", C.c_double), ("virial_ref", C.c_double), ("virial_pair", C.c_double)]


THERMO_DTYPE = np.dtype([("pe", "f8"), ("ke", "f8"), ("virial_ref", "f8"), ("virial_pair", "f8")])


class Stats(C.Structure):
    _fields_ = [("n_atoms", C.c_int64), ("n_ghost", C.c_int64), ("n_cells", C.c_int64 * 3),
                ("list_capacity", C.c_int64), ("max_neighbours", C.c_int64), ("capacity_growths", C.c_int64),
                ("n_builds", C.c_int64), ("n_steps", C.c_int64), ("n_launches", C.c_int64),
                ("device_bytes", C.c_int64), ("missing_type_pairs", C.c_int64)]

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "n_cells"}
        d["n_cells"] = list(self.n_cells)
        return d


class Nhc(C.Structure):
    """pisb_nhc == NHThermostatChain (src/ensemble/nvt.rs:4-17), 3 links."""
    _fields_ = [("chain_size", C.c_int32), ("pad", C.c_int32), ("start_temperature", C.c_double),
                ("end_temperature", C.c_double), ("target_temperature", C.c_double), ("xi", C.c_double * 3),
                ("eta", C.c_double * 3), ("g", C.c_double * 3), ("q", C.c_double * 3)]


class Mtk(C.Structure):
    """pisb_mtk == MTKBarostat (src/ensemble/npt.rs:11-22); matrices column-major."""
    _fields_ = [("target_pressure", C.c_double * 9), ("momentum", C.c_double * 9), ("w", C.c_double)]


class PisbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pisb error {code}: {msg}")
        self.code = code


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); the list is also what tests check against include/pisb200.h
SIGNATURES = {
    "pisb_version": (C.c_char_p, []),
    "pisb_device_count": (C.c_int, []),
    "pisb_create": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_double, C.POINTER(_vp)]),
    "pisb_destroy": (C.c_int, [_vp]),
    "pisb_last_error": (C.c_char_p, [_vp]),
    "pisb_set_box": (C.c_int, [_vp, _vp, _vp, _vp]),
    "pisb_upload": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "pisb_compute": (C.c_int, [_vp, C.c_int, _dp]),
    "pisb_step_nve": (C.c_int, [_vp, C.c_double, C.c_int64, _vp]),
    "pisb_download": (C.c_int, [_vp, _vp, _vp, _vp]),
    "pisb_download_begin": (C.c_int, [_vp, _vp, _vp, _vp]),
    "pisb_download_end": (C.c_int, [_vp]),
    "pisb_host_register": (C.c_int, [_vp, C.c_size_t]),
    "pisb_host_unregister": (C.c_int, [_vp]),
    "pisb_thermo_now": (C.c_int, [_vp, C.POINTER(Thermo)]),
    "pisb_start_velocities": (C.c_int, [_vp, C.c_double, C.c_uint64]),
    "pisb_verlet_step_nve_host": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_double, _dp]),
    "pisb_neighbours": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "pisb_invalidate_list": (C.c_int, [_vp]),
    "pisb_list_stats": (C.c_int, [_vp, _vp]),
    "pisb_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "pisb_set_profiling": (C.c_int, [_vp, C.c_int]),
    "pisb_timings": (C.c_int, [_vp, _vp, _vp]),
    "pisb_timings_reset": (C.c_int, [_vp]),
    "pisb_stream": (_vp, [_vp]),
    "pisb_synchronize": (C.c_int, [_vp]),
    "pisb_set_option": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "pisb_nhc_init": (C.c_int, [C.POINTER(Nhc), C.c_double, C.c_double, C.c_double]),
    "pisb_step_nvt_nhc": (C.c_int, [_vp, C.c_double, C.c_int64, C.POINTER(Nhc), C.c_int64, C.c_int64, _vp, _vp]),
    "pisb_mtk_init": (C.c_int, [C.POINTER(Mtk), _vp, C.c_double, C.c_int64, C.c_double]),
    "pisb_step_npt_mtk": (C.c_int, [_vp, C.c_double, C.c_int64, C.POINTER(Mtk), C.POINTER(Nhc), C.c_int64, C.c_int64, _vp,
                                    _vp, _vp]),
    "pisb_get_box": (C.c_int, [_vp, _vp, _vp]),
    "pisb_comm_unique_id": (C.c_int, [_vp, C.c_int]),
    "pisb_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "pisb_upload_owned": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "pisb_download_owned": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64)]),
    "pisb_download_owned_begin": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64)]),
    "pisb_owned_ids": (C.c_int, [_vp, C.c_int64, _vp, C.POINTER(C.c_int64)]),
}

_lib = None


def load():
    """Load libpisb200.so (building nothing: run __graft_entry__.build() first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                          "pis_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def check(handle, rc: int):
    if rc != PISB_OK:
        msg = load().pisb_last_error(handle).decode(errors="replace")
        raise PisbError(rc, msg)


# ---- pinned host arrays (for the end-to-end path) --------------------------------------------
def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked host memory (allocated through torch; plumbing only)."""
    import torch

    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    a = t.numpy().view(dtype).reshape(shape)
    a.flags.writeable = True
    _PINNED_KEEPALIVE[id(a)] = t
    return a


_PINNED_KEEPALIVE: dict = {}
