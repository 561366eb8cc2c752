"""The C-ABI library loads and exports every symbol include/pisb200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

from pis_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pisb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"PISB_API\s+[\w\s\*]+?\b(pisb_\w+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 24 and "pisb_step_nve" in syms and "pisb_comm_init" in syms
    lib = C.CDLL(capi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in pisb200.h but not exported"
        assert s in capi.SIGNATURES, f"{s} has no ctypes signature in pis_b200/capi.py"
    assert set(capi.SIGNATURES) <= set(syms)


def test_version_and_struct_layouts():
    lib = capi.load()
    assert b"sm_100a" in lib.pisb_version()
    assert C.sizeof(capi.Thermo) == 32 and capi.THERMO_DTYPE.itemsize == 32
    assert C.sizeof(capi.Stats) == 8 * 12
    assert lib.pisb_device_count() >= 0


def test_argument_validation_without_device():
    lib = capi.load()
    h = C.c_void_p()
    assert lib.pisb_create(0, 0, None, None, None, None, None, 1, 0.0, C.byref(h)) == capi.PISB_ERR_INVALID
    assert b"bad potential table" in lib.pisb_last_error(None)
    assert lib.pisb_destroy(None) == capi.PISB_OK
    assert lib.pisb_set_box(None, None, None, None) == capi.PISB_ERR_INVALID
