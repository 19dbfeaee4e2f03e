"""Size-independent checks at BASELINE.json's full sizes (no oracle at these sizes): in-core vs
integral-direct G agreement, symmetry, linearity.  usage: python tools/check_large.py alkane40 6-31G"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from fock_rs_b200 import _lib
from fock_rs_b200.fock import FockBuilder
from helpers import flat_for, random_symmetric_density

mol, basis = sys.argv[1], sys.argv[2]
fb = flat_for(mol, basis)
n = fb.n_basis
D = random_symmetric_density(n, seed=1)
with FockBuilder(fb, storage=_lib.STORAGE_DIRECT) as b:
    b.eri_build()
    t0 = time.time(); Gd = b.build_g(D); t1 = time.time()
    Gd2 = b.build_g(3.0 * D)
    print(f"direct: {1e3*(t1-t0):.1f} ms, timings {b.timings()['class_kernels_ms']:.2f} ms class kernels")
with FockBuilder(fb, storage=_lib.STORAGE_INCORE, keep_eri=False) as b:
    t0 = time.time(); b.eri_build(); t1 = time.time()
    Gi = b.build_g(D); t2 = time.time()
    tm = b.timings()
    print(f"in-core: build {1e3*(t1-t0):.1f} ms (class {tm['class_kernels_ms']:.2f}), G {1e3*(t2-t1):.2f} ms (device {tm['g_device_ms']:.3f})")
print("N", n, "max|G_incore - G_direct| =", float(np.max(np.abs(Gi - Gd))), " max|G| =", float(np.max(np.abs(Gd))))
print("symmetry", float(np.max(np.abs(Gd - Gd.T))), "linearity", float(np.max(np.abs(Gd2 - 3.0 * Gd))))
assert np.max(np.abs(Gi - Gd)) <= 1e-9 and np.max(np.abs(Gd2 - 3.0 * Gd)) <= 1e-9
print("LARGE_OK")
