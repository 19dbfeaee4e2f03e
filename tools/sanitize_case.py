#!/usr/bin/env python3
"""Small invocations of the hot path for compute-sanitizer (memcheck / racecheck / synccheck): every s/p/d class
kernel through both sinks, the K-split shuffles, the diagonal pass, the device sort/permutation, the digestion.
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py [sp|d]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fock_rs_b200 import _lib  # noqa: E402
from fock_rs_b200.fock import FockBuilder  # noqa: E402
from helpers import flat_for, random_symmetric_density  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "sp"
mol, basis = ("water_dimer", "6-31G") if which == "sp" else ("water", "6-31G**")
fb = flat_for(mol, basis)
D = random_symmetric_density(fb.n_basis, seed=1)
with FockBuilder(fb, storage=_lib.STORAGE_INCORE) as b:
    Gi = b.eri_build().build_g(D)
with FockBuilder(fb, storage=_lib.STORAGE_DIRECT) as b:
    Gd = b.eri_build().build_g(D)
with FockBuilder(fb, storage=_lib.STORAGE_DIRECT, deterministic=True) as b:
    Gx = b.eri_build().build_g(D)
print(f"sanitize_case {mol}/{basis}: N={fb.n_basis} |Gi-Gd|={np.max(np.abs(Gi - Gd)):.2e} |Gx-Gd|={np.max(np.abs(Gx - Gd)):.2e}")
