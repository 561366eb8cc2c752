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
    assert C.sizeof(capi.Stats) == 8 * 13
    assert lib.pisb_device_count() >= 0


def test_argument_validation_without_device():
    lib = capi.load()
    h = C.c_void_p()
    assert lib.pisb_create(0, 0, None, None, None, None, None, 1, 0.0, C.byref(h)) == capi.PISB_ERR_INVALID
    assert b"bad potential table" in lib.pisb_last_error(None)
    assert lib.pisb_destroy(None) == capi.PISB_OK
    assert lib.pisb_set_box(None, None, None, None) == capi.PISB_ERR_INVALID


def test_force_loops_have_no_local_memory_traffic():
    """The force loop is bound by the FP64 pipe and the L1TEX path; a register spill inside it (one STL + LDL per K-tile)
    costs ~8 % and is easy to pick up with an innocent-looking extra live value (it happened when the multi-GPU ghost test
    went into k_force_vv, and again when the guard-band fallback was a call inside the loop).  SASS check on the built
    library: every LOOP (backward branch) that holds the 256-bit neighbour gathers is found; the interior-warp loop (the
    first, ~85 % of the warps at 4M atoms) must be free of local-memory instructions, the minimum-image loop may keep at
    most one 32-bit spill pair.  (The rare guard-band fallback sits between the loops and passes its arguments on the
    stack: not counted.)"""
    import shutil
    import subprocess

    import pytest

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    kernels = {
        "k_force_v3<false>": "_ZN4pisb10k_force_v3ILb0EEEvNS_10Force2ArgsE",
        "k_force_vv<false,drift,single GPU>": "_ZN4pisb10k_force_vvILb0ELb1ELb0ELi512EEEvNS_11ForceVVArgsE",
        "k_force_vv<false,no drift,single GPU>": "_ZN4pisb10k_force_vvILb0ELb0ELb0ELi512EEEvNS_11ForceVVArgsE",
        "k_force_vv<false,drift,brick>": "_ZN4pisb10k_force_vvILb0ELb1ELb1ELi512EEEvNS_11ForceVVArgsE",
        "k_force_vv<false,no drift,brick>": "_ZN4pisb10k_force_vvILb0ELb0ELb1ELi512EEEvNS_11ForceVVArgsE",
        # four lanes per atom (the default above 75k atoms): one 32-bit reload per iteration of four tiles is tolerated
        "k_force_q<false,fused,drift,single GPU>": "_ZN4pisb9k_force_qILb0ELb1ELb1ELb0EEEvNS_11ForceVVArgsE",
        "k_force_q<false,fused,drift,brick>": "_ZN4pisb9k_force_qILb0ELb1ELb1ELb1EEEvNS_11ForceVVArgsE",
        "k_force_q<false,plain,single GPU>": "_ZN4pisb9k_force_qILb0ELb0ELb0ELb0EEEvNS_11ForceVVArgsE",
    }
    for name, sym in kernels.items():
        out = subprocess.run([cuobjdump, "-sass", "-fun", sym, capi.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
        ins = []
        for ln in out.splitlines():
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
            if m:
                ins.append((int(m.group(1), 16), m.group(2)))
        loops = []
        for addr, text in ins:
            m = re.search(r"\bBRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                body = [t for a_, t in ins if int(m.group(1), 16) <= a_ <= addr]
                if sum("LDG.E.ENL2.256" in t for t in body) >= 4:
                    loops.append(body)
        assert len(loops) >= 2, f"{name}: expected the interior and the minimum-image gather loops, found {len(loops)}"
        for k, body in enumerate(loops):
            local = [t for t in body if re.search(r"\b(STL|LDL)\b", t)]
            allowed = (1 if k == 0 else 2) if name.startswith("k_force_q") else (0 if k == 0 else 2)
            assert len(local) <= allowed, f"{name}: local-memory traffic inside gather loop {k}: {local[:4]}"
