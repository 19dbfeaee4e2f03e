"""HostStagedG (one PCIe round trip, device-side sum) against fock_build_g on one GPU (world = 1: no collective)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from helpers import flat_for, random_symmetric_density
from fock_rs_b200 import _lib, parallel
from fock_rs_b200.fock import FockBuilder

worst = 0.0
for storage in (_lib.STORAGE_DIRECT, _lib.STORAGE_INCORE):
    fb = flat_for("benzene", "6-31G")
    D = random_symmetric_density(fb.n_basis, seed=4)
    with FockBuilder(fb, storage=storage, keep_eri=False) as b:
        b.eri_build()
        want = b.build_g(D)
        st = parallel.HostStagedG(b)
        for _ in range(3):
            got = st(D)
        st2 = parallel.HostStagedG(b, torch.cuda.Stream())
        got2 = st2(2.0 * D)
        b.set_stream(None)
        again = b.build_g(D)
    worst = max(worst, float(np.abs(got - want).max()), float(np.abs(got2 - 2.0 * want).max()), float(np.abs(again - want).max()))
print("STAGED_OK" if worst <= 1e-11 else "STAGED_BAD", worst)
