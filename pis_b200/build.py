"""Build recipe for the PRODUCT's native pieces (run by __graft_entry__.build(); the test oracle has
its own recipe under oracle/).

  pis_b200/libpisb200.so   hand-written sm_100a kernels + C ABI (include/pisb200.h)      [product]
  pis_b200/pis_b200_cli    C++ host CLI mirroring the reference's `pis -i input.pis`       [product]

Everything is compiled in-tree with explicit nvcc / gcc command lines so the artefacts travel with
the repo snapshot to the GPU box (no JIT cache).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pis_b200", "csrc")
LIB = os.path.join(ROOT, "pis_b200", "libpisb200.so")
CLI = os.path.join(ROOT, "pis_b200", "pis_b200_cli")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # exact paths use __dmul_rn/__dadd_rn; -fmad=false makes "no contraction" the default everywhere
    "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-O2",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]}")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)


def _sources(dirpath: str, exts: tuple[str, ...]) -> list[str]:
    out = []
    for base, _dirs, files in os.walk(dirpath):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return sorted(out)


def build_library(force: bool = False, verbose_ptxas: bool = False) -> str:
    deps = _sources(CSRC, (".cu", ".cuh")) + [os.path.join(ROOT, "include", "pisb200.h")]
    if not force and _newer(LIB, deps):
        return LIB
    cus = [s for s in _sources(CSRC, (".cu",))]
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, *cus]
    if verbose_ptxas:
        cmd += ["-Xptxas", "-v"]
    _run(cmd)
    return LIB


def build_cli(force: bool = False) -> str | None:
    main = os.path.join(CSRC, "host", "main.cpp")
    if not os.path.exists(main):
        return None
    deps = _sources(os.path.join(CSRC, "host"), (".cpp", ".hpp")) + [LIB]
    if not force and _newer(CLI, deps):
        return CLI
    _run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
          "-I", os.path.join(CSRC, "host"), "-o", CLI, main, os.path.join(CSRC, "host", "pis_host.cpp"),
          "-L", os.path.dirname(LIB), "-lpisb200", "-Wl,-rpath,$ORIGIN"])
    return CLI


def build_all(force: bool = False) -> None:
    build_library(force)
    build_cli(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", LIB)
