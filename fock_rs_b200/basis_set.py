"""Host mirror of the reference's `basis` / `basis_set` data types.

Reference: basis/src/primitives.rs:7-48 (GaussianPrimitive, norm), basis/src/contracted_gaussian.rs:9-34
(ContractedGaussian), basis_set/src/lib.rs:15-118 (BSE-JSON schema + BasisSet::get), basis_set/src/atom.rs:7-44.

Same names, same argument meaning, same ordering of the produced basis functions (all shells
Cartesian, component order x-outer..z-inner => p = (z, y, x), d = (zz, yz, yy, xz, xy, xx);
primitives individually normalised, contracted functions not renormalised -- SURVEY.md F5).
`flatten()` produces the structure-of-arrays the C ABI takes (include/fock_b200.h: fock_basis).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "basis")

FRAC_2_PI = 0.6366197723675814  # std::f64::consts::FRAC_2_PI


def _powi(a: float, b: int) -> float:
    """f64::powi (compiler-rt __powidf2): square-and-multiply, reciprocal for negative exponents."""
    recip = b < 0
    b = abs(b)
    r = 1.0
    while True:
        if b & 1:
            r *= a
        b //= 2
        if b == 0:
            break
        a *= a
    return 1.0 / r if recip else r


@dataclass(frozen=True)
class GaussianPrimitive:
    """basis/src/primitives.rs:7-11; `coefficient` already carries the norm when built by `new`."""

    angular: tuple
    exponent: float
    coefficient: float

    @staticmethod
    def norm(angular: Sequence[int], alpha: float) -> float:
        """basis/src/primitives.rs:19-27."""
        i, j, k = angular
        prod = 1
        for l in (i, j, k):
            for m in range(l + 1, 2 * l + 1):
                prod *= m
        return math.sqrt(math.sqrt(_powi(FRAC_2_PI * alpha, 3))) * math.sqrt(_powi(8.0 * alpha, i + j + k) / float(prod))

    @classmethod
    def new(cls, angular, exponent: float, coefficient: float) -> "GaussianPrimitive":
        """basis/src/primitives.rs:41-48 (normalised)."""
        return cls(tuple(angular), exponent, coefficient * cls.norm(angular, exponent))

    @classmethod
    def new_unnormalized(cls, angular, exponent: float, coefficient: float) -> "GaussianPrimitive":
        return cls(tuple(angular), exponent, coefficient)


@dataclass
class ContractedGaussian:
    """basis/src/contracted_gaussian.rs:9-15: one basis FUNCTION (not a shell)."""

    position: tuple
    primitives: List[GaussianPrimitive]


@dataclass
class Atom:
    """basis_set/src/atom.rs:7-44.  `atom_type` is Z (the reference's enum discriminant)."""

    position: tuple
    atom_type: int
    basis: List[ContractedGaussian] = field(default_factory=list)

    def electron_count(self) -> int:
        return self.atom_type

    def point_charge(self):
        return (self.position, float(self.atom_type))


class BasisSet:
    """MolSSI BSE JSON basis set (basis_set/src/lib.rs:62-76) + `get` (:79-118)."""

    def __init__(self, doc: dict, general_contraction: bool = False):
        self.doc = doc
        # False (default) = the reference's `get`: function set `idx` of a shell uses `coefficients[idx]` with idx
        # enumerating `angular_momentum` (basis_set/src/lib.rs:86,107), so the extra columns of a generally
        # contracted shell (one angular momentum, several coefficient columns) are silently dropped.
        # True = SURVEY 8 row f2: every column of such a shell becomes its own contracted function set and
        # primitives whose coefficient is exactly zero (BSE pads general contractions with zeros) are left out.
        self.general_contraction = bool(general_contraction)
        self.name = doc.get("name", "?")
        self.elements = {int(z): e for z, e in doc["elements"].items()}

    @classmethod
    def from_file(cls, path: str, general_contraction: bool = False) -> "BasisSet":
        with open(path) as f:
            return cls(json.load(f), general_contraction)

    @classmethod
    def named(cls, name: str, general_contraction: bool = False) -> "BasisSet":
        """Shipped sets (trimmed copies of the reference's basis_set/src/data/*.json,
        basis_sets.rs:12-24): 'STO-3G', 'STO-6G', '3-21G', '6-31G', '6-31G*', '6-31G**', '6-311G'; plus
        'cc-pVDZ', a cc-pVDZ-SHAPED synthetic set for H and O (generally contracted; see its
        revision_description and SURVEY.md App. D) that needs `general_contraction=True` to load in full."""
        fname = name.replace("**", "_st_st").replace("*", "_st") + ".json"
        return cls.from_file(os.path.join(_DATA, fname), general_contraction)

    def get(self, position, atom_type: int) -> Atom:
        config = self.elements.get(int(atom_type))
        if config is None:
            raise KeyError("Failed to get atom from basis set")
        position = tuple(float(x) for x in position)
        functions: List[ContractedGaussian] = []
        for shell in config["electron_shells"]:
            columns = list(enumerate(shell["angular_momentum"]))
            general = self.general_contraction and len(shell["angular_momentum"]) == 1 and len(shell["coefficients"]) > 1
            if general:
                columns = [(idx, shell["angular_momentum"][0]) for idx in range(len(shell["coefficients"]))]
            for idx, angular in columns:
                angulars = [
                    (x, y, z)
                    for x in range(angular + 1)
                    for y in range(angular + 1)
                    for z in range(angular + 1)
                    if x + y + z == angular
                ]
                coeffs = shell["coefficients"][idx]
                for ang in angulars:
                    # zip() truncates to the shorter list exactly like Iterator::zip
                    prims = [GaussianPrimitive.new(ang, float(e), float(c)) for e, c in zip(shell["exponents"], coeffs)
                             if not (general and float(c) == 0.0)]
                    functions.append(ContractedGaussian(position, prims))
        return Atom(position, int(atom_type), functions)


@dataclass
class FlatBasis:
    """Structure-of-arrays form of `&[ContractedGaussian]` (include/fock_b200.h: fock_basis)."""

    center: np.ndarray  # [N,3] f64, bohr
    angular: np.ndarray  # [N,3] i32
    prim_offset: np.ndarray  # [N+1] i32
    exponent: np.ndarray  # [nprim] f64
    coefficient: np.ndarray  # [nprim] f64 (norm folded in)

    @property
    def n_basis(self) -> int:
        return int(self.center.shape[0])


def flatten(basis: Sequence[ContractedGaussian]) -> FlatBasis:
    """What the Rust shim does through the public accessors position()/primitives()/angular()/
    exponent()/coefficient() (basis/src/contracted_gaussian.rs:29-34, primitives.rs:77-89)."""
    n = len(basis)
    center = np.zeros((n, 3), dtype=np.float64)
    angular = np.zeros((n, 3), dtype=np.int32)
    off = np.zeros(n + 1, dtype=np.int32)
    exps, coefs = [], []
    for i, f in enumerate(basis):
        center[i] = f.position
        if not f.primitives:
            raise ValueError("contracted Gaussian without primitives")
        ang = f.primitives[0].angular
        for p in f.primitives:
            if tuple(p.angular) != tuple(ang):
                raise ValueError("primitives of one contracted Gaussian must share one angular triple")
            exps.append(p.exponent)
            coefs.append(p.coefficient)
        angular[i] = ang
        off[i + 1] = len(exps)
    return FlatBasis(center, angular, off, np.array(exps, dtype=np.float64), np.array(coefs, dtype=np.float64))


def molecule_basis(atoms: Sequence[Atom]) -> List[ContractedGaussian]:
    """scf/src/lib.rs:79-83: basis-function order = atom order, then shell order, then component order."""
    return [f for a in atoms for f in a.basis]
