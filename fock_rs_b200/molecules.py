"""Deterministic synthetic molecules for the BASELINE.json configs.

All generators return XYZ text in Angstrom; it goes through the same `parse_xyz` as shipped
files (so the reference's x1.89 conversion applies).
"""
from __future__ import annotations

import math
from typing import List, Tuple

_WATER = [("H", 0.7493682, 0.0, 0.4424329), ("O", 0.0, 0.0, -0.1653507), ("H", -0.7493682, 0.0, 0.4424329)]


def _xyz(atoms: List[Tuple[str, float, float, float]], title: str) -> str:
    lines = [str(len(atoms)), title]
    lines += [f"{s} {x:.10f} {y:.10f} {z:.10f}" for s, x, y, z in atoms]
    return "\n".join(lines) + "\n"


def water_cluster(nx: int, ny: int, nz: int, spacing: float = 3.0) -> str:
    """(H2O)_{nx*ny*nz}: un-rotated copies of water.xyz on a grid, copy (i,j,k) shifted by
    spacing*(i,j,k) Angstrom.  (H2O)16 = 4x2x2 (config 3), (H2O)32 = 4x4x2 (config 5)."""
    atoms = []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                for s, x, y, z in _WATER:
                    atoms.append((s, x + spacing * i, y + spacing * j, z + spacing * k))
    return _xyz(atoms, f"(H2O){nx * ny * nz} grid {nx}x{ny}x{nz} spacing {spacing} A")


def linear_alkane(n_carbon: int) -> str:
    """All-anti zig-zag C_nH_{2n+2} along x (config 4: n=40): C-C 1.54 A, angle CCC 112 deg,
    C-H 1.09 A, angle HCH 109.5 deg; atoms listed all C first, then H."""
    cc, ch = 1.54, 1.09
    half = math.radians(56.0)
    hh = math.radians(54.75)
    carbons = []
    for i in range(n_carbon):
        s = 1.0 if i % 2 == 1 else -1.0
        carbons.append((i * cc * math.sin(half), 0.0, s * 0.5 * cc * math.cos(half), s))
    atoms = [("C", x, y, z) for x, y, z, _ in carbons]
    for x, y, z, s in carbons:
        for sign in (1.0, -1.0):
            atoms.append(("H", x, sign * ch * math.sin(hh), z + s * ch * math.cos(hh)))
    # one extra in-plane H capping each end, continuing the zig-zag
    x0, _, z0, s0 = carbons[0]
    atoms.append(("H", x0 - ch * math.sin(half), 0.0, z0 - s0 * ch * math.cos(half)))
    x1, _, z1, s1 = carbons[-1]
    atoms.append(("H", x1 + ch * math.sin(half), 0.0, z1 - s1 * ch * math.cos(half)))
    return _xyz(atoms, f"C{n_carbon}H{2 * n_carbon + 2} all-anti")
