"""One fock_orbital_voxelize call on the renderer's grid (benzene/6-31G, 200^3) -- the ncu target for the voxel kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import flat_for
from fock_rs_b200.fock import FockBuilder

fb = flat_for(sys.argv[1] if len(sys.argv) > 1 else "benzene", sys.argv[2] if len(sys.argv) > 2 else "6-31G")
c = np.random.default_rng(20261017).uniform(-1.0, 1.0, fb.n_basis)
with FockBuilder(fb) as b:
    for _ in range(3):
        g = b.orbital_voxelize(c, 200, 20.0)
    print(b.orbital_stats(), float(np.abs(g).max()))
