"""Shared test inputs: the BASELINE.json molecules, flattened (tests only)."""
import functools

import numpy as np

from fock_rs_b200 import molecules
from fock_rs_b200.basis_set import BasisSet, flatten, molecule_basis
from fock_rs_b200.xyz import parse_xyz, read_xyz_file, shipped_molecule

_WATER = [("H", 0.7493682, 0.0, 0.4424329), ("O", 0.0, 0.0, -0.1653507), ("H", -0.7493682, 0.0, 0.4424329)]


def water_dimer_xyz(shift_y=3.0):
    return "\n".join(f"{s} {x} {y + dy} {z}" for dy in (0.0, shift_y) for s, x, y, z in _WATER)


@functools.lru_cache(maxsize=None)
def atoms_for(name, basis):
    bs = BasisSet.named(basis, general_contraction=basis.startswith("cc-"))
    if name == "water":
        return tuple(read_xyz_file(shipped_molecule("water"), bs))
    if name == "benzene":
        return tuple(read_xyz_file(shipped_molecule("benzene"), bs))
    if name == "water_dimer":
        return tuple(parse_xyz(water_dimer_xyz(3.0), bs))
    if name == "water_dimer6":
        return tuple(parse_xyz(water_dimer_xyz(6.0), bs))
    if name == "w16":
        return tuple(parse_xyz(molecules.water_cluster(4, 2, 2), bs))
    if name == "w4":
        return tuple(parse_xyz(molecules.water_cluster(2, 2, 1), bs))
    if name == "w32":
        return tuple(parse_xyz(molecules.water_cluster(4, 4, 2), bs))
    if name.startswith("alkane"):
        return tuple(parse_xyz(molecules.linear_alkane(int(name[6:])), bs))
    if name in ("ethane", "ethene", "ethyne", "ethanol", "cyclohexane"):
        return tuple(read_xyz_file(shipped_molecule(name), bs))
    raise KeyError(name)


def flat_for(name, basis):
    return flatten(molecule_basis(atoms_for(name, basis)))


def random_symmetric_density(n, seed=20261017):
    rng = np.random.default_rng(seed)
    d = rng.uniform(-1.0, 1.0, size=(n, n))
    return 0.5 * (d + d.T)
