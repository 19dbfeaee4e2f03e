"""Run under torchrun (one rank per GPU; 2, 4 or 8 ranks): sharded G builds against the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Checks, at the world size it is launched with:
  * small molecules, in-core (both shardings) and integral-direct, partial G summed by torch's all-reduce  vs the oracle's dense loop;
  * the same through the collective BEHIND the C ABI (fock_comm_init: fock_build_g returns the complete G on every rank);
  * BASELINE configs 3 and 4 at full size ((H2O)16 in-core + direct, C40H82 direct) and config 5 ((H2O)32 / 6-31G**) against the
    committed golden G matrices (tests/golden/g_*.npz, the oracle in integral-direct form);
  * deterministic mode: G bit-identical to the 1-rank result of the same build (fixed-point partials are summed as integers);
  * an SCF on the sharded engine (energy and iteration count of SURVEY.md App. B).
Prints MULTI_GPU_OK on success (exit code 1 otherwise).  `--quick` skips the full-size cases.
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
    from test_gpu_golden_large import CASES, load_golden

    quick = "--quick" in sys.argv
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    worst = 0.0
    ok = True

    def report(what, err, tol=1e-9):
        nonlocal worst, ok
        worst = max(worst, err)
        ok = ok and err <= tol
        if rank == 0:
            print(f"[world={world}] {what}: max|dG| = {err:.2e}" + ("" if err <= tol else "   <-- FAIL"), flush=True)

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
            report(f"{mol}/{basis} storage={storage} incore_shard={shard} (torch all-reduce)", float(np.max(np.abs(G - want))))
            with FockBuilder(fb, storage=storage, device=local, rank=rank, world=world, keep_eri=False, incore_shard=shard) as b:
                b.comm_init_torch()
                b.eri_build()
                G = b.build_g(D)  # complete on every rank
            report(f"{mol}/{basis} storage={storage} incore_shard={shard} (all-reduce inside fock_build_g)", float(np.max(np.abs(G - want))))
            # every rank must hold the same matrix
            t = torch.from_numpy(np.ascontiguousarray(G)).cuda()
            t0 = t.clone()
            dist.broadcast(t0, src=0)
            ok = ok and bool(torch.equal(t, t0))
        # deterministic mode: identical bits whatever the number of ranks
        with FockBuilder(fb, storage=_lib.STORAGE_DIRECT, device=local, deterministic=True) as b:
            G1 = b.eri_build().build_g(D)
        with FockBuilder(fb, storage=_lib.STORAGE_DIRECT, device=local, rank=rank, world=world, deterministic=True) as b:
            b.comm_init_torch()
            Gw = b.eri_build().build_g(D)
        same = bool(np.array_equal(G1, Gw))
        ok = ok and same
        report(f"{mol}/{basis} deterministic direct, {world} ranks vs 1 rank (bit-identical: {same})", float(np.max(np.abs(Gw - want))))

    if not quick:
        for tag, incore in (("w16_631g_faithful", True), ("alkane40_631g_faithful", False), ("w32_631gss_faithful", False)):
            mol, basis, mode, _, _ = CASES[tag]
            want, seed, _ = load_golden(tag)
            fb = flat_for(mol, basis)
            D = random_symmetric_density(fb.n_basis, seed=seed)
            with FockBuilder(fb, mode=mode, storage=_lib.STORAGE_DIRECT, device=local, rank=rank, world=world) as b:
                b.comm_init_torch()
                G = b.eri_build().build_g(D)
            report(f"{tag} direct, sharded over {world} ranks vs the golden oracle G", float(np.max(np.abs(G - want))))
            if incore:
                for shard in (0, 1):
                    with FockBuilder(fb, mode=mode, storage=_lib.STORAGE_INCORE, device=local, rank=rank, world=world, keep_eri=False, incore_shard=shard) as b:
                        b.comm_init_torch()
                        G = b.eri_build().build_g(D)
                    report(f"{tag} in-core (incore_shard={shard}), {world} ranks vs the golden oracle G", float(np.max(np.abs(G - want))))

    r = try_scf(atoms_for("water", "6-31G"), 100, 1e-6, 0, engine=FockEngine(device=local))
    e_err = abs(r.total_energy - (-76.0126589918))
    if rank == 0:
        print(f"SCF water/6-31G on {world} ranks: E = {r.total_energy:.10f} ({r.iterations} iterations), |dE| = {e_err:.1e}")
    ok = ok and e_err <= 2e-8 and r.iterations == 12
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(t[0]) != 1:
        if rank == 0:
            print("MULTI_GPU_FAILED")
        sys.exit(1)
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} worst max|dG| = {worst:.2e}")


if __name__ == "__main__":
    main()
