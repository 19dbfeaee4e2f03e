#!/usr/bin/env python3
"""Golden G matrices at BASELINE.json's full sizes (configs 3-5), from the CPU oracle in integral-direct form
(oracle_direct_g: the reference's tensor build basis/src/electron_tensor.rs:113-158 and G loop
scf/src/lib.rs:167-185 visited slot by slot, no N^4 storage), for the seeded density the GPU tests use.

The reference is Rust and cannot run here; the oracle is what pins the CUDA path at these sizes.  The jobs take
minutes to an hour of CPU (tabulated evaluator, all threads), which is why their results are committed as
fixtures:  tests/golden/g_<tag>.npz = { g_upper: upper triangle of G (row-major, i <= j), evaluated: integrals
evaluated incl. the diagonal ones, seed, flags, seconds, threads }.

usage: python tools/make_golden_large.py [tag ...]     (default: all that are missing)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import flat_for, random_symmetric_density  # noqa: E402

SEED = 20261017
# tag: (molecule, basis, oracle flags)   -- "accurate" = exact Boys + the (t-1) factor, reference cut kept
JOBS = {
    "w16_631g_faithful": ("w16", "6-31G", oracle.FAITHFUL),
    "w16_631g_accurate": ("w16", "6-31G", oracle.BOYS_EXACT | oracle.FIX_R),
    "alkane40_631g_faithful": ("alkane40", "6-31G", oracle.FAITHFUL),
    "w32_631gss_faithful": ("w32", "6-31G**", oracle.FAITHFUL),
    "w32_ccpvdz_faithful": ("w32", "cc-pVDZ", oracle.FAITHFUL),
    "w32_ccpvdz_accurate": ("w32", "cc-pVDZ", oracle.BOYS_EXACT | oracle.FIX_R),
}
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    tags = sys.argv[1:] or [t for t in JOBS if not os.path.exists(os.path.join(OUT, f"g_{t}.npz"))]
    for tag in tags:
        mol, basis, flags = JOBS[tag]
        fb = flat_for(mol, basis)
        ob = oracle.OracleBasis.from_flat(fb)
        n = fb.n_basis
        D = random_symmetric_density(n, seed=SEED)
        t0 = time.time()
        diag = oracle.diagonal(ob, flags, fast=True)
        G, nev = oracle.direct_g(ob, D, flags, fast=True, diag=diag)
        dt = time.time() - t0
        iu = np.triu_indices(n)
        np.savez_compressed(os.path.join(OUT, f"g_{tag}.npz"), g_upper=G[iu], evaluated=np.array([nev]), seed=np.array([SEED]),
                            flags=np.array([flags]), seconds=np.array([dt]), threads=np.array([oracle.num_threads()]),
                            sym_err=np.array([float(np.max(np.abs(G - G.T)))]))
        print(f"{tag}: N={n} evaluated={nev} {dt:.1f} s on {oracle.num_threads()} threads, |G-G^T|max={np.max(np.abs(G - G.T)):.2e}", flush=True)


if __name__ == "__main__":
    main()
