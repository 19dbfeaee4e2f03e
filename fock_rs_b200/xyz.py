"""Host mirror of chemfiles::xyz::read_xyz_file (reference: chemfiles/src/xyz.rs:13-32).

Angstrom -> bohr uses the reference's factor 1.89 (not 1.8897...).
"""
from __future__ import annotations

import os
import re
from typing import List

from .basis_set import Atom, BasisSet
from .periodic_table import from_symbol

ANGSTROM_TO_BOHR_REF = 1.89  # chemfiles/src/xyz.rs:23-26

_PATTERN = re.compile(r"\s*(\w+)\s+([+\-0-9.]+)\s+([+\-0-9.]+)\s+([0-9.+\-]+).*\b")
_MOLECULES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "molecules")


def parse_xyz(content: str, basis_set: BasisSet) -> List[Atom]:
    atoms = []
    for m in _PATTERN.finditer(content):
        x = float(m.group(2)) * ANGSTROM_TO_BOHR_REF
        y = float(m.group(3)) * ANGSTROM_TO_BOHR_REF
        z = float(m.group(4)) * ANGSTROM_TO_BOHR_REF
        atoms.append(basis_set.get((x, y, z), from_symbol(m.group(1))))
    return atoms


def read_xyz_file(path: str, basis_set: BasisSet) -> List[Atom]:
    with open(path) as f:
        return parse_xyz(f.read(), basis_set)


def shipped_molecule(name: str) -> str:
    """Path of a shipped geometry (copies of the reference's chemfiles/molecules/*.xyz)."""
    return os.path.join(_MOLECULES, name if name.endswith(".xyz") else name + ".xyz")
