"""GPU tests of the opt-in modes around the hot path: the Schwarz-screened mode (north_star: "opt-in screened mode"),
the deterministic fixed-point accumulation of the integral-direct G build, and bases with many distinct contraction
depths (more (class, primitive-count) sublists and jobs per class launch than any shipped set has)."""
import numpy as np
import pytest

import oracle
from fock_rs_b200 import _lib
from fock_rs_b200.basis_set import FlatBasis
from fock_rs_b200.fock import FockBuilder
from helpers import flat_for, random_symmetric_density

pytestmark = pytest.mark.gpu


def oracle_g(fb, D):
    """The CPU oracle's G (faithful arithmetic, reference cut) for the flat basis fb."""
    want, _ = oracle.direct_g(oracle.OracleBasis.from_flat(fb), D, oracle.FAITHFUL)
    return want


@pytest.mark.parametrize("storage", [_lib.STORAGE_INCORE, _lib.STORAGE_DIRECT], ids=["incore", "direct"])
def test_schwarz_tau_1e6_is_the_reference_cut(storage):
    """FOCK_SCREEN_SCHWARZ applies the reference's own estimate sqrt((ij|ij)(kl|kl)) > tau
    (basis/src/electron_tensor.rs:129-134) with a caller-chosen tau: tau = 1e-6 must reproduce the reference cut
    bit for bit (same integrals kept, same arithmetic)."""
    fb = flat_for("water_dimer6", "STO-3G")  # 3844 of 5460 off-diagonal integrals are cut
    D = random_symmetric_density(fb.n_basis, seed=9)
    with FockBuilder(fb, storage=storage, screen=_lib.SCREEN_REFERENCE) as b:
        b.eri_build()
        G_ref = b.build_g(D)
        V_ref = b.eri_export_packed() if storage == _lib.STORAGE_INCORE else None
    with FockBuilder(fb, storage=storage, screen=_lib.SCREEN_SCHWARZ, tau=1e-6) as b:
        b.eri_build()
        G_tau = b.build_g(D)
        V_tau = b.eri_export_packed() if storage == _lib.STORAGE_INCORE else None
    if V_ref is not None:
        assert np.array_equal(V_ref, V_tau)
        assert np.count_nonzero(V_ref == 0.0) >= 3844
    assert np.max(np.abs(G_ref - G_tau)) <= 1e-12


@pytest.mark.parametrize("mol,basis", [("water_dimer", "STO-3G"), ("w4", "6-31G**")])
def test_schwarz_tau_bounds_the_error_in_g(mol, basis):
    """Every integral the screen drops has |(ij|kl)| <= sqrt((ij|ij)(kl|kl)) <= tau, and each enters G through
    at most 6 slots with weights <= 2|D|: for |D| <= 1,  max|G_tau - G_full| <= 4 * tau * (dropped integrals)."""
    fb = flat_for(mol, basis)
    n = fb.n_basis
    D = random_symmetric_density(n, seed=10)
    ob = oracle.OracleBasis.from_flat(fb)
    diag = oracle.diagonal(ob, oracle.FAITHFUL, fast=True)
    with FockBuilder(fb, storage=_lib.STORAGE_DIRECT, screen=_lib.SCREEN_OFF) as b:
        G_full = b.eri_build().build_g(D)
    prev = None
    for tau in (1e-4, 1e-7, 1e-10):
        with FockBuilder(fb, storage=_lib.STORAGE_DIRECT, screen=_lib.SCREEN_SCHWARZ, tau=tau) as b:
            G_tau = b.eri_build().build_g(D)
        est = np.sqrt(np.abs(np.outer(diag, diag)))
        dropped = int((np.count_nonzero(est <= tau) - np.count_nonzero(np.diag(est) <= tau)) // 2)
        err = float(np.max(np.abs(G_tau - G_full)))
        assert err <= 4.0 * tau * max(dropped, 1) + 1e-12
        if prev is not None:
            assert err <= prev + 1e-12  # tightening tau never makes G worse
        prev = err
    assert prev <= 1e-8  # tau = 1e-10 is converged to the unscreened matrix


@pytest.mark.parametrize("mol,basis,mode", [("w4", "6-31G", _lib.MODE_FAITHFUL), ("w4", "6-31G**", _lib.MODE_ACCURATE)])
def test_deterministic_direct_g_is_bit_identical_and_right(mol, basis, mode):
    """opts.deterministic: 64-bit fixed-point accumulation (integer atomics are associative).  G must be bit-identical
    from run to run, agree with the FP64-atomic build to rounding, and with the oracle to the north_star tolerance."""
    fb = flat_for(mol, basis)
    n = fb.n_basis
    D = random_symmetric_density(n, seed=12)
    ob = oracle.OracleBasis.from_flat(fb)
    oflags = oracle.FAITHFUL if mode == _lib.MODE_FAITHFUL else (oracle.BOYS_EXACT | oracle.FIX_R)
    want, _ = oracle.direct_g(ob, D, oflags)
    runs = []
    for _ in range(2):
        with FockBuilder(fb, mode=mode, storage=_lib.STORAGE_DIRECT, deterministic=True) as b:
            b.eri_build()
            runs.append(b.build_g(D))
            runs.append(b.build_g(D))
    for g in runs[1:]:
        assert np.array_equal(g, runs[0])
    with FockBuilder(fb, mode=mode, storage=_lib.STORAGE_DIRECT) as b:
        G_fp = b.eri_build().build_g(D)
    assert np.max(np.abs(runs[0] - G_fp)) <= 1e-11
    assert np.max(np.abs(runs[0] - want)) <= 1e-9
    assert np.max(np.abs(runs[0] - runs[0].T)) == 0.0
    # a density 1e6 times larger rescales the fixed point instead of overflowing it
    with FockBuilder(fb, mode=mode, storage=_lib.STORAGE_DIRECT, deterministic=True) as b:
        G_big = b.eri_build().build_g(1e6 * D)
    assert np.max(np.abs(G_big - 1e6 * runs[0])) <= 1e-9 * 1e6


def test_deterministic_mode_is_refused_for_incore_storage():
    fb = flat_for("water", "STO-3G")
    with pytest.raises(_lib.FockError) as e:
        FockBuilder(fb, storage=_lib.STORAGE_INCORE, deterministic=True)
    assert e.value.code == _lib.ERR_UNSUPPORTED


def many_depth_basis(n_centres=2, seed=3):
    """s and p shells with 1..6 primitives on two centres: 6 contraction depths give 18 distinct K = Ka*Kb in the ss
    pair class, i.e. 171 (bra sublist, ket sublist) jobs in the (ss|ss) launch (the old limit was 64)."""
    rng = np.random.default_rng(seed)
    center, angular, off, exps, coefs = [], [], [0], [], []
    for a in range(n_centres):
        pos = np.array([1.9 * a, 0.4 * (a % 2), 0.25 * a])
        for depth in range(1, 7):
            e = np.sort(rng.uniform(0.15, 6.0, depth))[::-1] * (1.0 + 0.1 * a)
            for l, comps in ((0, [(0, 0, 0)]), (1, [(0, 0, 1), (0, 1, 0), (1, 0, 0)])):
                if l == 1 and depth % 2 == 0:
                    continue
                c = rng.uniform(0.2, 1.0, depth)
                for lmn in comps:
                    center.append(pos)
                    angular.append(lmn)
                    for k in range(depth):
                        exps.append(e[k])
                        coefs.append(c[k] * oracle.norm(lmn, e[k]))
                    off.append(len(exps))
    return FlatBasis(np.array(center), np.array(angular, dtype=np.int32), np.array(off, dtype=np.int32), np.array(exps), np.array(coefs))


def test_many_contraction_depths_do_not_overflow_the_job_tables():
    """a basis with >= 5 distinct contraction depths in one angular class used to fail fock_eri_build
    ('too many sublist pairs in one class launch').  Tensor, both G paths, against the oracle."""
    fb = many_depth_basis()
    n = fb.n_basis
    ob = oracle.OracleBasis.from_flat(fb)
    want, _ = oracle.from_basis_packed(ob, oracle.FAITHFUL, fast=True)
    D = random_symmetric_density(n, seed=13)
    G_want = oracle.build_g(want, n, D)
    with FockBuilder(fb) as b:
        got = b.eri_build().eri_export_packed()
        G_in = b.build_g(D)
        st = b.stats()
    assert st["partial_shells"] == 0 and st["general_shell_quartets_evaluated"] == 0
    assert np.max(np.abs(got - want)) <= 1e-10
    assert np.max(np.abs(G_in - G_want)) <= 1e-9
    with FockBuilder(fb, storage=_lib.STORAGE_DIRECT) as b:
        G_d = b.eri_build().build_g(D)
    assert np.max(np.abs(G_d - G_want)) <= 1e-9


def test_d_shells_are_never_fragmented():
    """the regrouping tolerance must not split legitimate Cartesian d shells (fock_stats[12])."""
    for mol, basis in (("water", "6-31G**"), ("w4", "6-31G**"), ("water", "cc-pVDZ")):
        with FockBuilder(flat_for(mol, basis)) as b:
            assert b.stats()["partial_shells"] == 0


@pytest.mark.parametrize("storage", [_lib.STORAGE_INCORE, _lib.STORAGE_DIRECT], ids=["incore", "direct"])
def test_page_locked_matrices_take_the_same_path(storage):
    """fock_build_g with D and G in page-locked memory (fock_pinned_alloc, ABI v5) copies straight to / from the device; the result
    is the one the staged path gives for pageable buffers, and mixing the two kinds works."""
    fb = flat_for("water_dimer", "6-31G")
    D = random_symmetric_density(fb.n_basis, seed=5)
    with FockBuilder(fb, storage=storage) as b:
        b.eri_build()
        G_pageable = b.build_g(D)
        Dp, Gp = b.pinned_matrix(), b.pinned_matrix()
        Dp[...] = D
        Gp[...] = np.nan
        assert b.build_g(Dp, out=Gp) is Gp
        G_mixed = b.build_g(Dp)  # pinned in, pageable out
        want = oracle_g(fb, D)
        for G in (Gp, G_mixed):
            assert np.max(np.abs(G - G_pageable)) <= 1e-13  # FP64 atomics: same sums, order differs from run to run
            assert np.max(np.abs(G - want)) <= 1e-9
        with pytest.raises(ValueError):
            b.build_g(Dp, out=np.empty((fb.n_basis, fb.n_basis)))  # row-major out


def test_spatial_and_diagonal_pair_orders_agree(monkeypatch):
    """The pair order of the sublists (spatial order + CSR of the surviving quartets by default; FOCK_PAIR_ORDER=diag = sorted by
    the largest diagonal integral, survivors of a row are a prefix) only changes which lane evaluates which quartet: both
    enumerate exactly the quartets the reference's cut keeps (same counts) and give the oracle's G."""
    fb = flat_for("w4", "6-31G**")
    D = random_symmetric_density(fb.n_basis, seed=11)
    want = oracle_g(fb, D)
    res = {}
    for order in ("spatial", "diag"):
        monkeypatch.setenv("FOCK_PAIR_ORDER", order)
        for bits in (("0", "4") if order == "spatial" else ("4",)):
            monkeypatch.setenv("FOCK_ORDER_BUCKET_BITS", bits)
            with FockBuilder(fb, storage=_lib.STORAGE_DIRECT) as b:
                b.eri_build()
                G = b.build_g(D)
                res[(order, bits)] = (G, b.stats()["shell_quartets_evaluated"])
    counts = {v[1] for v in res.values()}
    assert len(counts) == 1, res.keys()
    for G, _ in res.values():
        assert np.max(np.abs(G - want)) <= 1e-9


def test_incore_high_l_register_kernels_match_the_warp_per_quartet_fallback(monkeypatch):
    """In-core builds of the class pairs with total angular momentum >= 6 ((dp|dp) ... (dd|dd)) run register-resident slice
    kernels (dclass_hit_*.cu); FOCK_NO_HI_INCORE=1 sends them through the warp-per-quartet kernel as r01 did.  Same packed
    tensor (1e-12: two evaluation schemes of the same sums) and the oracle's G from both."""
    fb = flat_for("ethene", "6-31G**")  # two d shells: every class pair up to (dd|dd) between different shells
    D = random_symmetric_density(fb.n_basis, seed=13)
    want = oracle_g(fb, D)
    out = {}
    for key, env in (("register", None), ("fallback", "1")):
        if env is None:
            monkeypatch.delenv("FOCK_NO_HI_INCORE", raising=False)
        else:
            monkeypatch.setenv("FOCK_NO_HI_INCORE", env)
        with FockBuilder(fb, storage=_lib.STORAGE_INCORE, keep_eri=True) as b:
            b.eri_build()
            out[key] = (b.eri_export_packed(), b.build_g(D), b.stats()["launches_build"])
    assert out["register"][2] > out["fallback"][2]  # the slice launches are there
    assert np.max(np.abs(out["register"][0] - out["fallback"][0])) <= 1e-12
    for key in out:
        assert np.max(np.abs(out[key][1] - want)) <= 1e-9


@pytest.mark.parametrize("storage", [_lib.STORAGE_INCORE, _lib.STORAGE_DIRECT], ids=["incore", "direct"])
@pytest.mark.parametrize("xyz,basis", [("H 0 0 0", "STO-3G"), ("H 0 0 0\nH 0 0 50.0", "STO-3G"), ("H 0 0 0\nH 0 0 50.0", "6-31G**"),
                                        ("O 0 0 0", "6-31G**")],
                         ids=["one-function", "two-far-atoms", "two-far-atoms-p", "one-atom-spd"])
def test_degenerate_plans(xyz, basis, storage):
    """Plans at the edges of the CSR construction: a single basis function (one pair, one quartet, which is a diagonal one),
    two atoms 50 A apart (every cross pair falls under the cut: rows without survivors, jobs without tiles), one atom with s, p and d
    shells (every class pair present, all on one centre)."""
    from fock_rs_b200.basis_set import BasisSet, flatten, molecule_basis
    from fock_rs_b200.xyz import parse_xyz

    fb = flatten(molecule_basis(tuple(parse_xyz(xyz, BasisSet.named(basis)))))
    D = random_symmetric_density(fb.n_basis, seed=3)
    want = oracle_g(fb, D)
    with FockBuilder(fb, storage=storage) as b:
        b.eri_build()
        G = b.build_g(D)
    assert np.max(np.abs(G - want)) <= 1e-9
    with FockBuilder(fb, storage=storage, screen=_lib.SCREEN_SCHWARZ, tau=0.5) as b:  # a cut that leaves (almost) nothing but the diagonal
        b.eri_build()
        G_cut = b.build_g(D)
    assert np.all(np.isfinite(G_cut)) and np.max(np.abs(G_cut - G_cut.T)) <= 1e-12
