"""The reference's shipped example (example/input.pis + example/argon4000.txt, BASELINE configs[0]) for tests that run
where /root/reference does not exist.  The data file is regenerated (byte-identical: tests/golden/make_example_inputs.py,
test_host.py::test_example_inputs_are_the_reference_bytes), the script is the committed byte copy."""
from __future__ import annotations

import os
from decimal import Decimal

from pis_b200.lattice import fcc_positions

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rust_display(x: float) -> str:
    """Rust `{}` formatting of an f64: shortest round-trip digits, never scientific, 0.0 -> "0"."""
    if x != x:
        return "NaN"
    s = format(Decimal(repr(float(x))), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s if s not in ("", "-") else "0"


def argon4000_text() -> str:
    """example/argon4000.txt: 10^3 FCC cells of a = 5.41 in sapphire's atom order (src/bin/sapphire/main.rs:61-77)."""
    pos = fcc_positions(5.41, 10, 10, 10)
    lines = ["", "", "4000 atoms", "1 atom types", "", "0.0 54.1 xlo xhi", "0.0 54.1 ylo yhi", "0.0 54.1 zlo zhi", "",
             "Masses", "1 39.948", "", "PairCoeffs", "1 0.238 3.405 8.5", "", "Atoms"]
    lines += [f"{i + 1} 1 {rust_display(p[0])} {rust_display(p[1])} {rust_display(p[2])}" for i, p in enumerate(pos)]
    return "\n".join(lines) + "\n"


def example_script(nve: bool = False, steps: int | None = None) -> str:
    """example/input.pis as shipped (NPT); nve=True removes its `fix ... npt` line; steps rewrites `run`."""
    with open(os.path.join(GOLDEN, "example_input.pis"), "rb") as fh:
        text = fh.read().decode("utf-8")
    out = []
    for ln in text.split("\n"):
        tok = ln.split()
        if nve and tok[:1] == ["fix"]:
            continue
        if steps is not None and tok[:1] == ["run"]:
            ln = f"run {steps}"
        out.append(ln)
    return "\n".join(out)


def write_example(dirpath, nve: bool = False, steps: int | None = None) -> str:
    """Lay out <dir>/input.pis and <dir>/example/argon4000.txt the way the reference's repository does."""
    os.makedirs(os.path.join(dirpath, "example"), exist_ok=True)
    with open(os.path.join(dirpath, "example", "argon4000.txt"), "w") as fh:
        fh.write(argon4000_text())
    path = os.path.join(dirpath, "input.pis")
    with open(path, "w", encoding="utf-8") as fh:
        fh.write(example_script(nve, steps))
    return path
