"""Run under torchrun (one rank per GPU): sharded in-core and direct G builds vs the CPU oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import oracle
    from fock_rs_b200 import _lib, parallel
    from fock_rs_b200.fock import FockBuilder, FockEngine
    from fock_rs_b200.scf import try_scf
    from helpers import atoms_for, flat_for, random_symmetric_density

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    worst = 0.0
    for mol, basis in [("water_dimer", "STO-3G"), ("benzene", "6-31G"), ("water", "6-31G**")]:
        fb = flat_for(mol, basis)
        n = fb.n_basis
        ob = oracle.OracleBasis.from_flat(fb)
        packed, _ = oracle.from_basis_packed(ob, oracle.FAITHFUL, fast=True)
        D = random_symmetric_density(n, seed=4)
        want = oracle.build_g(packed, n, D)
        for storage, shard in ((_lib.STORAGE_INCORE, 0), (_lib.STORAGE_INCORE, 1), (_lib.STORAGE_DIRECT, 0)):
            with FockBuilder(fb, storage=storage, device=local, rank=rank, world=world, keep_eri=False, incore_shard=shard) as b:
                b.eri_build()
                G = parallel.allreduce_host_matrix(b.build_g(D))
            err = float(np.max(np.abs(G - want)))
            worst = max(worst, err)
            if rank == 0:
                print(f"{mol}/{basis} storage={storage} incore_shard={shard} world={world}: max|dG| = {err:.2e}")
    r = try_scf(atoms_for("water", "6-31G"), 100, 1e-6, 0, engine=FockEngine(device=local))
    e_err = abs(r.total_energy - (-76.0126589918))
    if rank == 0:
        print(f"SCF water/6-31G on {world} ranks: E = {r.total_energy:.10f} ({r.iterations} iterations), |dE| = {e_err:.1e}")
    ok = worst <= 1e-9 and e_err <= 2e-8 and r.iterations == 12
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(t[0]) != 1:
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_OK")


if __name__ == "__main__":
    main()
