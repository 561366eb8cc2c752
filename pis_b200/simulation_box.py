"""SimulationBox -- host mirror of src/simulation_box.rs (h, cached h_inv, pbc).

The device never recomputes h_inv: it is produced here, with nalgebra's 3x3 `try_inverse`
algorithm (adjugate / determinant; nalgebra 0.34.1 linalg/inverse.rs), and handed to the kernels.
"""
from __future__ import annotations

import numpy as np


def _det3(m) -> float:
    # nalgebra Matrix3::determinant (m[r][c], 0-based)
    minor_m12_m23 = m[1][1] * m[2][2] - m[2][1] * m[1][2]
    minor_m11_m23 = m[1][0] * m[2][2] - m[2][0] * m[1][2]
    minor_m11_m22 = m[1][0] * m[2][1] - m[2][0] * m[1][1]
    return m[0][0] * minor_m12_m23 - m[0][1] * minor_m11_m23 + m[0][2] * minor_m11_m22


def _try_inverse3(m):
    m11, m12, m13 = m[0]
    m21, m22, m23 = m[1]
    m31, m32, m33 = m[2]
    minor_m12_m23 = m22 * m33 - m32 * m23
    minor_m11_m23 = m21 * m33 - m31 * m23
    minor_m11_m22 = m21 * m32 - m31 * m22
    det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22
    if det == 0.0:
        return None
    return [
        [minor_m12_m23 / det, (m13 * m32 - m33 * m12) / det, (m12 * m23 - m22 * m13) / det],
        [-minor_m11_m23 / det, (m11 * m33 - m31 * m13) / det, (m13 * m21 - m23 * m11) / det],
        [minor_m11_m22 / det, (m12 * m31 - m32 * m11) / det, (m11 * m22 - m21 * m12) / det],
    ]


class SimulationBox:
    """ref: src/simulation_box.rs:5-15.  `h`/`h_inv` are 3x3 numpy arrays indexed [row, col]."""

    def __init__(self, h, pbc=(True, True, True)):
        hm = [[float(h[r][c]) for c in range(3)] for r in range(3)]
        inv = _try_inverse3(hm)
        if inv is None:
            raise ValueError("Box matrix should be invertible")  # simulation_box.rs:13 panics
        self.h = np.array(hm, dtype=np.float64)
        self.h_inv = np.array(inv, dtype=np.float64)
        self.pbc = [bool(p) for p in pbc]

    @classmethod
    def from_lammps_data(cls, xlo, xhi, ylo, yhi, zlo, zhi, xy=0.0, xz=0.0, yz=0.0):
        """ref: src/simulation_box.rs:44-65 (columns a, b, c; pbc always true)."""
        ax, ay, az = xhi - xlo, yhi - ylo, zhi - zlo
        h = [[ax, xy, xz], [0.0, ay, yz], [0.0, 0.0, az]]
        return cls(h, (True, True, True))

    def volume(self) -> float:
        """ref: src/simulation_box.rs:67-69"""
        return abs(_det3(self.h.tolist()))

    # column-major flat views for the C ABI
    def h_colmajor(self):
        return np.ascontiguousarray(self.h.T.reshape(9))

    def h_inv_colmajor(self):
        return np.ascontiguousarray(self.h_inv.T.reshape(9))
