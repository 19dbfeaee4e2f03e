#!/usr/bin/env python3
"""One build + one G build of a named molecule/basis on cuda:0, timed after a warm-up pass; prints the library's own
timings and counters.  The quick A/B probe used by tools/gpu_variants.sh (TAG labels the line, FOCK_B200_LIB picks an
experimental build).

usage: tools/time_config.py <molecule> <basis> <screen: 0 ref | 1 off> <storage: 0 in-core | 1 direct>
  e.g. tools/time_config.py alkane40 6-31G 0 1      tools/time_config.py w32 'cc-pVDZ' 0 1
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fock_rs_b200.fock import FockBuilder  # noqa: E402
from helpers import flat_for, random_symmetric_density  # noqa: E402


def main():
    mol, basis, screen, storage = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    fb = flat_for(mol, basis)
    D = random_symmetric_density(fb.n_basis)
    # RANK / WORLD emulate one rank's share of a multi-GPU run on a single GPU (the tiles are dealt round-robin inside the kernels;
    # G is then a partial sum): per-rank kernel time against the 1-GPU time / WORLD = what the tails and fixed costs take
    rank, world = int(os.environ.get("RANK_EMU", "0")), int(os.environ.get("WORLD_EMU", "1"))
    with FockBuilder(fb, screen=screen, storage=storage, keep_eri=False, rank=rank, world=world) as b:
        b.eri_build()
        b.build_g(D)  # warm-up
        t0 = time.time()
        b.eri_build()
        t1 = time.time()
        b.build_g(D)
        t2 = time.time()
        tm, st = b.timings(), b.stats()
        reps = []  # a few more G builds: min / median of the class-kernel time (direct) or the digestion (in-core)
        for _ in range(int(os.environ.get("REPS", "6"))):
            if storage == 0:
                b.eri_build()
            b.build_g(D)
            t = b.timings()
            reps.append(t["class_kernels_ms"])
        reps.sort()
    tflops = st["model_flops"] / (tm["class_kernels_ms"] * 1e-3) * 1e-12 if tm["class_kernels_ms"] > 0 else 0.0
    print(os.environ.get("TAG", ""), ("rank %d/%d" % (rank, world)) if world > 1 else "", mol, "screen", screen, "storage", storage,
          "build %.2f ms g %.2f ms | class %.3f (min %.3f med %.3f of %d) g_dev %.3f sort %.2f plan %.2f tiles %d TF %.2f"
          % (1e3 * (t1 - t0), 1e3 * (t2 - t1), tm["class_kernels_ms"], reps[0] if reps else 0.0, reps[len(reps) // 2] if reps else 0.0, len(reps), tm["g_device_ms"], tm["host_sort_ms"], tm["host_plan_ms"],
             st["warp_tiles"], tflops), flush=True)
    print({k: st[k] for k in ("shells", "shell_quartets_unique", "shell_quartets_evaluated", "general_shell_quartets_evaluated",
                              "prim_quartets_evaluated", "launches_build", "launches_g")})


if __name__ == "__main__":
    main()
