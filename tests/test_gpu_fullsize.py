"""Parity at BASELINE.json's full sizes, where the CPU oracle would take minutes to hours: the
size-independent properties the domain offers -- the integral-direct and the in-core G builds are two
independent digestions of the same integrals and must agree; G is symmetric and linear in D; the
reference's cut only ever removes magnitude below what its estimate allows."""
import numpy as np
import pytest

from fock_rs_b200 import _lib
from fock_rs_b200.fock import FockBuilder
from helpers import flat_for, random_symmetric_density

pytestmark = pytest.mark.gpu


def _gpu_mem_gb():
    import torch

    return torch.cuda.get_device_properties(0).total_memory / 2**30


@pytest.mark.parametrize("mol,basis,need_gb", [("w16", "6-31G", 8), ("alkane40", "6-31G", 100)])
def test_incore_and_direct_digestions_agree(mol, basis, need_gb):
    if _gpu_mem_gb() < need_gb:
        pytest.skip("not enough device memory for the packed supermatrix")
    fb = flat_for(mol, basis)
    n = fb.n_basis
    D = random_symmetric_density(n, seed=1)
    with FockBuilder(fb, storage=_lib.STORAGE_DIRECT) as b:
        b.eri_build()
        Gd = b.build_g(D)
        Gd3 = b.build_g(3.0 * D)
    with FockBuilder(fb, storage=_lib.STORAGE_INCORE, keep_eri=False) as b:
        Gi = b.eri_build().build_g(D)
    assert np.max(np.abs(Gi - Gd)) <= 1e-9
    assert np.max(np.abs(Gd - Gd.T)) == 0.0
    assert np.max(np.abs(Gd3 - 3.0 * Gd)) <= 1e-9


def test_reference_cut_is_bounded_by_its_estimate_w16():
    # every integral the 1e-6 cut removes has |(ij|kl)| <= sqrt((ij|ij)(kl|kl)) <= 1e-6 (Schwarz), so
    # for a density with |D| <= 1 the two G matrices differ by at most (number of cut integrals) * 3e-6
    fb = flat_for("w16", "6-31G")
    D = random_symmetric_density(fb.n_basis, seed=2)
    with FockBuilder(fb, storage=_lib.STORAGE_DIRECT, screen=_lib.SCREEN_REFERENCE) as b:
        G_cut = b.eri_build().build_g(D)
    with FockBuilder(fb, storage=_lib.STORAGE_DIRECT, screen=_lib.SCREEN_OFF) as b:
        G_all = b.eri_build().build_g(D)
    diff = float(np.max(np.abs(G_cut - G_all)))
    assert 0.0 < diff < 5e-3  # the cut is visible (F3) and small


def test_config5_d_shells_direct_runs_and_is_linear():
    fb = flat_for("w32", "6-31G**")
    assert fb.n_basis == 800
    D = random_symmetric_density(800, seed=3)
    with FockBuilder(fb, storage=_lib.STORAGE_DIRECT) as b:
        b.eri_build()
        G = b.build_g(D)
        st = b.stats()
    # d shells run through the class engine (register-resident slice kernels up to total L = 5, the
    # warp-per-quartet fallback above), not through the general shell-quartet list
    assert st["general_shell_quartets_evaluated"] == 0 and st["shell_quartets_evaluated"] > 0
    assert np.all(np.isfinite(G)) and np.max(np.abs(G - G.T)) == 0.0
    # trace identity: sum_ij D_ij G_ij = sum_ijkl D_ij D_kl [(ij|kl) - 1/2 (ik|jl)] is invariant under D -> -D
    with FockBuilder(fb, storage=_lib.STORAGE_DIRECT) as b:
        Gm = b.eri_build().build_g(-D)
    assert np.max(np.abs(G + Gm)) <= 1e-9


def test_d_shell_classes_under_the_cut_direct_vs_incore_vs_oracle():
    """(H2O)4 / 6-31G** (100 bf): every d class pair, screened staircases, K-split jobs; the in-core tensor
    against the oracle's, and G through both storage modes against the oracle's loop."""
    import oracle

    fb = flat_for("w4", "6-31G**")
    assert fb.n_basis == 100
    D = random_symmetric_density(100, seed=11)
    ob = oracle.OracleBasis.from_flat(fb)
    want, _ = oracle.from_basis_packed(ob, oracle.FAITHFUL, fast=True)
    with FockBuilder(fb) as b:
        got = b.eri_build().eri_export_packed()
        G_incore = b.build_g(D)
    assert np.max(np.abs(got - want)) <= 1e-10
    with FockBuilder(fb, storage=_lib.STORAGE_DIRECT) as b:
        G_direct = b.eri_build().build_g(D)
    G_want = oracle.build_g(want, 100, D)
    assert np.max(np.abs(G_incore - G_want)) <= 1e-9
    assert np.max(np.abs(G_direct - G_want)) <= 1e-9
    assert np.max(np.abs(G_direct - G_direct.T)) == 0.0
