"""Pins the CPU oracle (oracle/fock_oracle.c) before anything is compared against it.

Level 0: the reference's own known-answer test on this path -- the `hermite_expansion` doctest
(basis/src/utils.rs:24), bit-exact.  Level 1: scalar values (Boys / ERI / S / T) and SCF
energies recorded in SURVEY.md App. B from an independent restatement of the reference.
"""
import math

import numpy as np
import pytest

import oracle
from fock_rs_b200.scf import nuclear_repulsion, try_scf
from helpers import atoms_for, flat_for
from oracle_engine import OracleEngine


def test_hermite_expansion_doctest_bit_exact():
    # basis/src/utils.rs:24: assert_eq!(hermite_expansion([1, 2, 3], 2.5, 3.0, 5.0), 1.9872458203327215e-9)
    assert oracle.hermite_expansion([1, 2, 3], 2.5, 3.0, 5.0) == 1.9872458203327215e-9


def test_hermite_expansion_out_of_range_is_zero():
    assert oracle.hermite_expansion([1, 1, 3], 0.4, 1.0, 2.0) == 0.0
    assert oracle.hermite_expansion([1, 1, -1], 0.4, 1.0, 2.0) == 0.0


BOYS_GOLDEN = [
    (0.1, 0, 9.67643312635591779e-01),
    (0.1, 10, 4.35578042613670585e-02),
    (1.0, 0, 7.46824132875526736e-01),
    (1.0, 4, 4.96232415472480856e-02),
    (7.3, 2, 4.56128341913639677e-03),
    (12.345, 8, 3.32884773499983128e-06),
    (29.9, 0, 1.64945695857447838e-01),
    (29.99, 5, 2.06915361345973396e-07),
    (30.0, 0, 1.61802159379641597e-01),
    (30.0, 10, 1.75199731737021577e-10),
    (100.0, 3, 1.66167548522392123e-07),
    (1e-17, 3, 1.0 / 7.0),
]


@pytest.mark.parametrize("t,n,want", BOYS_GOLDEN)
def test_boys_faithful_golden(t, n, want):
    assert oracle.boys(t, n) == pytest.approx(want, rel=4e-16, abs=0)


def test_boys_faithful_quirks():
    # F1: the LUT+downward recursion is visibly wrong near T=30 (exact F0(29.9)=0.16207)
    assert abs(oracle.boys(29.9, 0) - oracle.boys_exact(29.9, 0)) > 2e-3
    # order > 10 panics in the reference (boys/src/lib.rs:59-62) -> NaN here
    assert math.isnan(oracle.boys(1.0, 11))


def test_boys_exact_against_closed_forms():
    # F0(T) = sqrt(pi/T)/2 * erf(sqrt(T))
    for t in (1e-3, 0.5, 3.0, 12.0, 29.9, 45.0, 80.0):
        want = 0.5 * math.sqrt(math.pi / t) * math.erf(math.sqrt(t))
        assert oracle.boys_exact(t, 0) == pytest.approx(want, rel=2e-14)
    # downward recursion identity F_{n-1} = (2T F_n + e^-T)/(2n-1)
    for t in (0.3, 7.0, 33.0):
        for n in range(1, 9):
            lhs = oracle.boys_exact(t, n - 1)
            rhs = (2 * t * oracle.boys_exact(t, n) + math.exp(-t)) / (2 * n - 1)
            assert lhs == pytest.approx(rhs, rel=1e-13)


def test_coulomb_auxiliary_f2_quirk():
    # x branch lacks the (t-1) factor (basis/src/utils.rs:98-105): R_300 differs, R_030 / R_003 do not
    d = np.array([0.3, 0.3, 0.3])
    dsq = float(d @ d)
    rx = oracle.coulomb_auxiliary(3, 0, 0, 0, 1.1, d, dsq)
    ry = oracle.coulomb_auxiliary(0, 3, 0, 0, 1.1, d, dsq)
    rz = oracle.coulomb_auxiliary(0, 0, 3, 0, 1.1, d, dsq)
    assert ry == pytest.approx(rz, rel=1e-14)
    assert abs(rx - ry) > 1e-3
    assert oracle.coulomb_auxiliary(3, 0, 0, 0, 1.1, d, dsq, oracle.FIX_R) == pytest.approx(ry, rel=1e-14)


def test_water_sto3g_layout_and_folded_coefficients():
    fb = flat_for("water", "STO-3G")
    assert fb.n_basis == 7
    assert fb.angular.tolist() == [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0], [0, 0, 0]]
    np.testing.assert_allclose(fb.center[0], [1.416305898, 0.0, 0.836198181], atol=1e-9)
    np.testing.assert_allclose(fb.center[1], [0.0, 0.0, -0.312512823], atol=1e-9)
    np.testing.assert_allclose(fb.coefficient[:3], [0.276934355089, 0.267838851619, 0.0834736711328], rtol=1e-11)
    np.testing.assert_allclose(fb.coefficient[fb.prim_offset[3] : fb.prim_offset[4]], [1.6754501961, 1.053568044, 0.166902879086], rtol=1e-10)


ERI_GOLDEN = [
    ((0, 0, 0, 0), 7.746059443199189e-01),
    ((1, 1, 1, 1), 4.785065752279990e00),
    ((3, 3, 3, 3), 8.801590896119943e-01),
    ((5, 5, 5, 5), 7.852702008902424e-01),
    ((5, 5, 3, 3), 7.852702008902424e-01),
    ((5, 0, 5, 0), 1.041235984266916e-01),
    ((5, 1, 5, 6), 1.244359525117720e-02),
    ((6, 0, 2, 1), 3.800022242915398e-02),
]


@pytest.mark.parametrize("q,want", ERI_GOLDEN)
def test_eri_faithful_golden(q, want):
    ob = oracle.OracleBasis.from_flat(flat_for("water", "STO-3G"))
    assert oracle.eri(ob, *q) == pytest.approx(want, rel=2e-15)
    assert oracle.eri(ob, *q, fast=True) == pytest.approx(want, rel=2e-15)


def test_eri_modes_golden():
    ob = oracle.OracleBasis.from_flat(flat_for("water", "STO-3G"))
    assert oracle.eri(ob, 5, 5, 5, 5, oracle.FIX_R) == pytest.approx(8.801590896119943e-01, rel=2e-15)
    assert oracle.eri(ob, 5, 0, 5, 0, oracle.ACCURATE) == pytest.approx(1.041235896877646e-01, rel=1e-13)
    assert oracle.eri(ob, 5, 1, 5, 6, oracle.ACCURATE) == pytest.approx(1.244351782462679e-02, rel=1e-13)
    assert oracle.eri(ob, 6, 0, 2, 1, oracle.ACCURATE) == pytest.approx(3.799981720246087e-02, rel=1e-13)


@pytest.mark.parametrize("flags", [oracle.FAITHFUL, oracle.FIX_R, oracle.ACCURATE])
def test_fast_evaluator_matches_recursive(flags):
    # d shells included (6-31G** water): total L up to 8
    fb = flat_for("water", "6-31G**")
    ob = oracle.OracleBasis.from_flat(fb)
    rng = np.random.default_rng(7)
    worst = 0.0
    for _ in range(120):
        a, b, c, d = (int(x) for x in rng.integers(0, fb.n_basis, 4))
        slow = oracle.eri(ob, a, b, c, d, flags)
        fast = oracle.eri(ob, a, b, c, d, flags, fast=True)
        worst = max(worst, abs(slow - fast))
    assert worst <= 1e-13


def test_one_electron_golden():
    atoms = atoms_for("water", "STO-3G")
    ob = oracle.OracleBasis.from_flat(flat_for("water", "STO-3G"))
    pos = np.array([a.position for a in atoms])
    chg = np.array([float(a.atom_type) for a in atoms])
    S, T, V = oracle.one_electron(ob, pos, chg)
    assert S[0, 1] == pytest.approx(0.05298725473887029, rel=1e-14)
    assert T[1, 1] == pytest.approx(29.003204066085097, rel=1e-14)
    assert nuclear_repulsion(atoms) == pytest.approx(9.1269609703, abs=1e-9)
    np.testing.assert_allclose(np.diag(S), 1.0, atol=1e-6)


def test_packed_index_is_eightfold_canonical():
    n = 5
    seen = {}
    for x in range(n):
        for y in range(n):
            for z in range(n):
                for w in range(n):
                    k = oracle.packed_index(x, y, z, w)
                    assert k == oracle.packed_index(y, x, z, w) == oracle.packed_index(x, y, w, z) == oracle.packed_index(z, w, x, y)
                    seen[k] = 1
    assert len(seen) == oracle.packed_len(n) == 120


SCF_GOLDEN = [
    ("water", "STO-3G", oracle.FAITHFUL, -74.9874357034, 7),
    ("water", "STO-3G", oracle.FIX_R, -74.9660902656, 7),
    ("water", "STO-3G", oracle.ACCURATE, -74.9644720469, 7),
    ("water", "6-31G", oracle.FAITHFUL, -76.0126589918, 12),
    ("water", "6-31G", oracle.ACCURATE, -75.9826350360, 12),
]


@pytest.mark.parametrize("mol,basis,flags,energy,iters", SCF_GOLDEN)
def test_scf_golden_energies(mol, basis, flags, energy, iters):
    r = try_scf(atoms_for(mol, basis), 100, 1e-6, 0, engine=OracleEngine(flags, fast=False))
    assert r is not None
    assert r.total_energy == pytest.approx(energy, abs=2e-9)
    assert r.iterations == iters


def test_reference_cut_water_dimer():
    # F3: the reference zeroes integrals whose estimate sqrt((ij|ij)(kl|kl)) is not > 1e-6
    atoms = atoms_for("water_dimer", "STO-3G")
    cut = OracleEngine(oracle.FAITHFUL, fast=False)
    r_cut = try_scf(atoms, 100, 1e-6, 0, engine=cut)
    nocut = OracleEngine(oracle.NO_CUT, fast=False)
    r_nocut = try_scf(atoms, 100, 1e-6, 0, engine=nocut)
    assert nocut.n_evaluated == 5460 and cut.n_evaluated == 5460 - 1135
    assert r_cut.total_energy == pytest.approx(-149.908188846054, abs=2e-9)
    assert r_nocut.total_energy == pytest.approx(-149.908190247774, abs=2e-9)
    assert r_cut.iterations == r_nocut.iterations == 8
    assert np.max(np.abs(cut.packed - nocut.packed)) == pytest.approx(9.4e-7, rel=0.05)


def test_build_g_matches_dense_einsum():
    fb = flat_for("water", "STO-3G")
    ob = oracle.OracleBasis.from_flat(fb)
    packed, _ = oracle.from_basis_packed(ob, oracle.FAITHFUL, fast=True)
    n = fb.n_basis
    dense = np.zeros((n,) * 4)
    for i in range(n):
        for j in range(n):
            for k in range(n):
                for l in range(n):
                    dense[i, j, k, l] = packed[oracle.packed_index(i, j, k, l)]
    rng = np.random.default_rng(3)
    D = rng.uniform(-1, 1, (n, n))
    D = 0.5 * (D + D.T)
    G = oracle.build_g(packed, n, D)
    want = np.einsum("kl,ijkl->ij", D, dense) - 0.5 * np.einsum("kl,ikjl->ij", D, dense)
    np.testing.assert_allclose(G, want, atol=1e-13)
