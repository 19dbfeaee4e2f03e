"""End-to-end RHF timing with the device-resident SCF loop (accurate mode: the faithful-mode trajectory
is chaotic beyond tiny molecules).  usage: python tools/scf_timing.py"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from fock_rs_b200 import _lib
from fock_rs_b200.scf import try_scf_device
from helpers import atoms_for

for mol, basis, storage in [("water", "6-31G", 0), ("benzene", "6-31G", 0), ("w16", "6-31G", 0), ("w16", "6-31G", 1), ("alkane40", "6-31G", 1),
                            ("w16", "6-31G**", 0), ("w32", "6-31G**", 1)]:  # d shells: in-core at 400 bf (25.7 GB supermatrix), direct at 800 bf
    atoms = atoms_for(mol, basis)
    with torch.cuda.stream(torch.cuda.Stream()):
        try_scf_device(atoms_for("water", "STO-3G"), 50, 1e-6, 0)  # warm up cuSOLVER/cuBLAS handles
        for rep in range(2):  # the first run of a size pays cuSOLVER/cuBLAS workspace and allocator growth
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = try_scf_device(atoms, 100, 1e-6, 0, mode=_lib.MODE_ACCURATE, screen=_lib.SCREEN_REFERENCE, storage=storage)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
    print(f"{mol}/{basis} N={r.n_basis} storage={'direct' if storage else 'in-core'}: E = {r.total_energy:.10f} Eh, {r.iterations} iterations, "
          f"{dt*1e3:.1f} ms total ({dt*1e3/(r.iterations+1):.2f} ms per iteration incl. setup)", flush=True)
