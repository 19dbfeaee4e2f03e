#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the CPU oracle (the reference itself is Rust
and cannot run here; the oracle is pinned against the reference's doctest and the survey's probe values
in tests/test_oracle.py).  Re-run only when the oracle changes:  python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import flat_for, random_symmetric_density  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

for mol, basis, tag in [("water", "STO-3G", "water_sto3g"), ("water", "6-31G", "water_631g"), ("water_dimer", "STO-3G", "water_dimer_sto3g")]:
    fb = flat_for(mol, basis)
    ob = oracle.OracleBasis.from_flat(fb)
    D = random_symmetric_density(fb.n_basis, seed=20261017)
    data = {"density": D}
    for name, flags in [("faithful", oracle.FAITHFUL), ("accurate", oracle.ACCURATE)]:
        packed, nev = oracle.from_basis_packed(ob, flags, fast=False)  # the un-memoised recursion
        data[f"packed_{name}"] = packed
        data[f"g_{name}"] = oracle.build_g(packed, fb.n_basis, D, parallel=False)
        data[f"evaluated_{name}"] = np.array([nev])
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **data)

t = np.concatenate([np.linspace(0.0, 35.0, 701), [1e-17, 29.9999999, 30.0, 100.0]])
boys = np.array([[oracle.boys(float(x), n) for x in t] for n in range(11)])
boys_exact = np.array([[oracle.boys_exact(float(x), n) for x in t] for n in range(11)])
np.savez_compressed(os.path.join(OUT, "boys.npz"), t=t, faithful=boys, exact=boys_exact)
print("wrote", sorted(os.listdir(OUT)))

# Orbital evaluation (section 8 row f4): synthetic coefficient columns, random points, one small voxel grid.
from test_orbital import CASES as ORB_CASES, coeff_for, points_for  # noqa: E402

orb = {}
pts = points_for(flat_for("water", "6-31G"), 300)
orb["points"] = pts
for mol, basis in ORB_CASES:
    tag = f"{mol}_{basis}".replace("*", "s").replace("-", "")
    fb = flat_for(mol, basis)
    ob = oracle.OracleBasis.from_flat(fb)
    c = coeff_for(fb.n_basis)
    orb[f"coeff_{tag}"] = c
    orb[f"values_{tag}"] = oracle.orbital_eval(ob, c, pts)
    if tag == "water_631G":
        orb["grid16_water_631G"] = oracle.voxelize(ob, c, 16, 20.0)
np.savez_compressed(os.path.join(OUT, "orbital.npz"), **orb)
print("wrote orbital.npz")
