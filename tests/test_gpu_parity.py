"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same inputs.  Tolerances are the north_star's: ERI <= 1e-10 Eh absolute, G <= 1e-9, converged RHF
energy <= 1e-8 Eh with the same iteration count.
"""
import numpy as np
import pytest

import oracle
from fock_rs_b200 import _lib
from fock_rs_b200.fock import ElectronTensor, FockBuilder, FockEngine, electron_repulsion_int
from fock_rs_b200.basis_set import molecule_basis
from fock_rs_b200.scf import try_scf
from helpers import atoms_for, flat_for, random_symmetric_density
from oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu

ERI_TOL = 1e-10
G_TOL = 1e-9
E_TOL = 1e-8

MODES = [
    (_lib.MODE_FAITHFUL, oracle.FAITHFUL),
    (_lib.MODE_FIX_R, oracle.FIX_R),
    (_lib.MODE_BOYS_ACCURATE, oracle.BOYS_EXACT),
    (_lib.MODE_ACCURATE, oracle.BOYS_EXACT | oracle.FIX_R),
]


def assert_same_cut(got, want):
    """The reference's cut must zero the same slots on both sides.  (Integrals that vanish by
    symmetry may be an exact 0.0 on one side and ~1e-17 on the other, hence the 1e-12 floor.)"""
    assert not np.any((want == 0.0) & (np.abs(got) > 1e-12))
    assert not np.any((got == 0.0) & (np.abs(want) > 1e-12))


def oracle_flags(mode_flags, screen):
    return mode_flags | (oracle.NO_CUT if screen == _lib.SCREEN_OFF else 0)


@pytest.mark.parametrize("mode,oflags", MODES)
def test_generic_kernel_random_quartets_with_d_shells(mode, oflags):
    fb = flat_for("water", "6-31G**")  # s, p and Cartesian d functions: total L up to 8
    ob = oracle.OracleBasis.from_flat(fb)
    rng = np.random.default_rng(11)
    q = rng.integers(0, fb.n_basis, size=(400, 4)).astype(np.int32)
    with FockBuilder(fb, mode=mode) as b:
        got = b.eri_batch(q)
    want = np.array([oracle.eri(ob, *map(int, r), oflags, fast=True) for r in q])
    assert np.max(np.abs(got - want)) <= ERI_TOL


def test_electron_repulsion_int_single_call():
    basis = molecule_basis(atoms_for("water", "STO-3G"))
    ob = oracle.OracleBasis.from_flat(flat_for("water", "STO-3G"))
    got = electron_repulsion_int(basis[5], basis[1], basis[5], basis[6])
    assert got == pytest.approx(oracle.eri(ob, 5, 1, 5, 6), abs=1e-13)
    assert got == pytest.approx(1.244359525117720e-02, abs=1e-13)


@pytest.mark.parametrize("kernel", [0, 1], ids=["class", "generic"])
@pytest.mark.parametrize("mode,oflags", MODES)
@pytest.mark.parametrize("mol,basis", [("water", "STO-3G"), ("water", "6-31G"), ("water_dimer", "STO-3G")])
def test_packed_tensor_matches_oracle(mol, basis, mode, oflags, kernel):
    fb = flat_for(mol, basis)
    ob = oracle.OracleBasis.from_flat(fb)
    want, _ = oracle.from_basis_packed(ob, oflags, fast=True)
    with FockBuilder(fb, mode=mode, kernel=kernel) as b:
        got = b.eri_build().eri_export_packed()
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= ERI_TOL
    assert_same_cut(got, want)


def test_reference_cut_counts_water_dimer():
    fb = flat_for("water_dimer", "STO-3G")
    with FockBuilder(fb) as b:
        cut = b.eri_build().eri_export_packed()
    with FockBuilder(fb, screen=_lib.SCREEN_OFF) as b:
        nocut = b.eri_build().eri_export_packed()
    n = fb.n_basis
    npair = n * (n + 1) // 2
    # 1135 of the 5460 off-diagonal integrals fall under the cut (many of them vanish
    # by symmetry anyway, so only the ones that are visibly nonzero without the cut are counted here)
    assert int(np.sum((cut == 0.0) & (np.abs(nocut) > 1e-12))) > 300
    ob = oracle.OracleBasis.from_flat(fb)
    want_cut, nev = oracle.from_basis_packed(ob, oracle.FAITHFUL, fast=True)
    assert nev == 5460 - 1135
    assert_same_cut(cut, want_cut)
    assert np.max(np.abs(cut - nocut)) == pytest.approx(9.4e-7, rel=0.05)
    assert npair * (npair + 1) // 2 == cut.size


def test_index_operator_is_eightfold_symmetric():
    fb = flat_for("water", "STO-3G")
    ob = oracle.OracleBasis.from_flat(fb)
    t = ElectronTensor(FockBuilder(fb).eri_build())
    v = t[(5, 1, 6, 5)]
    for idx in [(1, 5, 6, 5), (5, 1, 5, 6), (6, 5, 5, 1), (5, 6, 1, 5)]:
        assert t[idx] == v
    assert v == pytest.approx(oracle.eri(ob, 5, 1, 5, 6), abs=ERI_TOL)
    t.builder.close()


@pytest.mark.parametrize("mode,oflags,screen", [(_lib.MODE_FAITHFUL, oracle.FAITHFUL, _lib.SCREEN_REFERENCE), (_lib.MODE_ACCURATE, oracle.BOYS_EXACT | oracle.FIX_R, _lib.SCREEN_OFF)])
def test_benzene_631g_tensor_and_g(mode, oflags, screen):
    fb = flat_for("benzene", "6-31G")
    assert fb.n_basis == 66
    ob = oracle.OracleBasis.from_flat(fb)
    want, nev = oracle.from_basis_packed(ob, oracle_flags(oflags, screen), fast=True)
    D = random_symmetric_density(fb.n_basis)
    with FockBuilder(fb, mode=mode, screen=screen) as b:
        b.eri_build()
        got = b.eri_export_packed()
        G = b.build_g(D)
        st = b.stats()
    assert np.max(np.abs(got - want)) <= ERI_TOL
    assert_same_cut(got, want)
    if screen == _lib.SCREEN_REFERENCE:
        assert 2443155 - nev == 383121  # integrals zeroed by the reference's cut
    G_want = oracle.build_g(want, fb.n_basis, D)
    assert np.max(np.abs(G - G_want)) <= G_TOL
    assert np.max(np.abs(G - G.T)) == 0.0
    assert st["shell_quartets_evaluated"] > 0


@pytest.mark.parametrize("mol,basis", [("water", "STO-3G"), ("water", "6-31G**"), ("ethene", "6-31G")])
def test_g_build_fixed_density(mol, basis):
    fb = flat_for(mol, basis)
    ob = oracle.OracleBasis.from_flat(fb)
    packed, _ = oracle.from_basis_packed(ob, oracle.FAITHFUL, fast=True)
    D = random_symmetric_density(fb.n_basis, seed=5)
    with FockBuilder(fb) as b:
        G = b.eri_build().build_g(D)
        G2 = b.build_g(2.0 * D)
    assert np.max(np.abs(G - oracle.build_g(packed, fb.n_basis, D))) <= G_TOL
    np.testing.assert_allclose(G2, 2.0 * G, rtol=0, atol=1e-11)  # linearity


@pytest.mark.parametrize("mode,oflags", [(_lib.MODE_FAITHFUL, oracle.FAITHFUL), (_lib.MODE_ACCURATE, oracle.BOYS_EXACT | oracle.FIX_R)])
def test_one_electron_matrices(mode, oflags):
    atoms = atoms_for("water", "6-31G**")
    fb = flat_for("water", "6-31G**")
    ob = oracle.OracleBasis.from_flat(fb)
    pos = np.array([a.position for a in atoms])
    chg = np.array([float(a.atom_type) for a in atoms])
    S0, T0, V0 = oracle.one_electron(ob, pos, chg, oflags)
    with FockBuilder(fb, mode=mode) as b:
        S, T, V = b.one_electron(pos, chg)
    assert np.max(np.abs(S - S0)) <= 1e-12
    assert np.max(np.abs(T - T0)) <= 1e-11
    assert np.max(np.abs(V - V0)) <= 1e-10


SCF_CASES = [
    ("water", "STO-3G", _lib.MODE_FAITHFUL, _lib.SCREEN_REFERENCE, oracle.FAITHFUL, -74.9874357034, 7),
    ("water", "STO-3G", _lib.MODE_ACCURATE, _lib.SCREEN_OFF, oracle.ACCURATE, -74.9644720469, 7),
    ("water", "6-31G", _lib.MODE_FAITHFUL, _lib.SCREEN_REFERENCE, oracle.FAITHFUL, -76.0126589918, 12),
    ("water_dimer", "STO-3G", _lib.MODE_FAITHFUL, _lib.SCREEN_REFERENCE, oracle.FAITHFUL, -149.908188846054, 8),
]


@pytest.mark.parametrize("mol,basis,mode,screen,oflags,energy,iters", SCF_CASES)
def test_scf_energy_and_iteration_count(mol, basis, mode, screen, oflags, energy, iters):
    atoms = atoms_for(mol, basis)
    r = try_scf(atoms, 100, 1e-6, 0, engine=FockEngine(mode=mode, screen=screen))
    ro = try_scf(atoms, 100, 1e-6, 0, engine=OracleEngine(oflags, fast=True))
    assert r is not None and ro is not None
    assert r.iterations == ro.iterations == iters
    assert abs(r.total_energy - ro.total_energy) <= E_TOL
    assert abs(r.total_energy - energy) <= 2e-8  # SURVEY.md App. B (10 decimals recorded)


def test_benzene_accurate_scf_energy():
    # faithful-mode SCF is chaotic beyond tiny molecules: energy parity for config 2 is
    # asserted in accurate mode, GPU-accurate vs oracle-accurate.
    atoms = atoms_for("benzene", "6-31G")
    r = try_scf(atoms, 100, 1e-6, 0, engine=FockEngine(mode=_lib.MODE_ACCURATE, screen=_lib.SCREEN_OFF))
    assert r is not None
    assert r.iterations == 10
    assert abs(r.total_energy - (-230.6222568588)) <= 2e-8


def test_error_codes():
    fb = flat_for("water", "STO-3G")
    with FockBuilder(fb) as b:
        with pytest.raises(_lib.FockError) as ei:
            b.build_g(np.eye(fb.n_basis))
        assert ei.value.code == _lib.ERR_STATE
        with pytest.raises(_lib.FockError) as ei:
            b.eri_batch([[0, 0, 0, 99]])
        assert ei.value.code == _lib.ERR_BAD_ARG


DIRECT_CASES = [
    ("water", "STO-3G", _lib.MODE_FAITHFUL, _lib.SCREEN_REFERENCE, oracle.FAITHFUL),
    ("water_dimer", "STO-3G", _lib.MODE_FAITHFUL, _lib.SCREEN_REFERENCE, oracle.FAITHFUL),
    ("water", "6-31G**", _lib.MODE_FAITHFUL, _lib.SCREEN_REFERENCE, oracle.FAITHFUL),  # d functions: generic remainder
    ("ethene", "6-31G", _lib.MODE_ACCURATE, _lib.SCREEN_OFF, oracle.ACCURATE),
    ("benzene", "6-31G", _lib.MODE_FAITHFUL, _lib.SCREEN_REFERENCE, oracle.FAITHFUL),
]


@pytest.mark.parametrize("mol,basis,mode,screen,oflags", DIRECT_CASES)
def test_integral_direct_g_matches_oracle_and_incore(mol, basis, mode, screen, oflags):
    fb = flat_for(mol, basis)
    ob = oracle.OracleBasis.from_flat(fb)
    packed, _ = oracle.from_basis_packed(ob, oflags, fast=True)
    D = random_symmetric_density(fb.n_basis, seed=9)
    G_want = oracle.build_g(packed, fb.n_basis, D)
    with FockBuilder(fb, mode=mode, screen=screen, storage=_lib.STORAGE_DIRECT) as b:
        b.eri_build()
        G_direct = b.build_g(D)
        G_again = b.build_g(D)
    with FockBuilder(fb, mode=mode, screen=screen, keep_eri=False) as b:
        G_incore = b.eri_build().build_g(D)
    assert np.max(np.abs(G_direct - G_want)) <= G_TOL
    assert np.max(np.abs(G_incore - G_want)) <= G_TOL
    assert np.max(np.abs(G_direct - G_again)) <= 1e-12  # atomics reorder sums: run-to-run jitter only
    assert np.max(np.abs(G_direct - G_direct.T)) == 0.0


def test_direct_mode_scf_water():
    atoms = atoms_for("water", "6-31G")
    r = try_scf(atoms, 100, 1e-6, 0, engine=FockEngine(storage=_lib.STORAGE_DIRECT))
    assert r is not None and r.iterations == 12
    assert abs(r.total_energy - (-76.0126589918)) <= 2e-8


def test_supermatrix_scatter_equals_gather_definition():
    # kernel=1 forms P by the gather definition from V; kernel=0 scatters integrals as produced
    fb = flat_for("water_dimer", "STO-3G")
    D = random_symmetric_density(fb.n_basis, seed=2)
    with FockBuilder(fb, kernel=0) as b:
        g0 = b.eri_build().build_g(D)
    with FockBuilder(fb, kernel=1) as b:
        g1 = b.eri_build().build_g(D)
    assert np.max(np.abs(g0 - g1)) <= 1e-12


def test_two_rank_sharding_matches_oracle():
    """World-size-2 run (one process per GPU, NCCL): tiles sharded, supermatrix reduce-scattered,
    partial G all-reduced.  Needs two GPUs; the gloo test in test_host_cpu.py covers the host logic."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os

    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_check.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_GPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("mode,exact", [(_lib.MODE_FAITHFUL, False), (_lib.MODE_ACCURATE, True)])
def test_device_boys_function_matches_reference(mode, exact):
    """boys::boys(t, n) on the device (the per-interval tables of the class kernels for n <= 4, the
    generic evaluator above) against the oracle on a dense grid incl. the T = 30 switch-over."""
    rng = np.random.default_rng(3)
    t = np.concatenate([rng.uniform(0.0, 36.0, 20000), np.arange(0, 200) * 0.15, np.arange(1, 200) * 0.15 * (1 - 1e-16),
                        [0.0, 1e-17, 1e-9, 29.99, 29.999999, 30.0, 30.0000001, 45.0, 100.0, 250.0]])
    with FockBuilder(flat_for("water", "STO-3G"), mode=mode) as b:
        for n in range(0, 11):
            got = b.boys(t, n)
            want = np.array([(oracle.boys_exact if exact else oracle.boys)(float(x), n) for x in t])
            assert np.max(np.abs(got - want)) <= 2e-15, n
            # relative check where the value is not negligible; 1e-9 because the accurate mode switches to the
            # asymptotic form at T = 40, dropping terms of order exp(-40) (absolute 1e-19)
            big = np.abs(want) > 1e-12
            assert np.max(np.abs(got - want)[big] / np.abs(want)[big]) <= 1e-9, n
            small_t = big & (t < 29.0)
            assert np.max(np.abs(got - want)[small_t] / np.abs(want)[small_t]) <= 1e-13, n
        with pytest.raises(_lib.FockError) as ei:
            b.boys(t[:4], 11)
        assert ei.value.code == _lib.ERR_BOYS_ORDER


@pytest.mark.parametrize("mode,oflags", MODES)
@pytest.mark.parametrize("mol,basis", [("water", "6-31G**"), ("ethene", "6-31G**")])
@pytest.mark.parametrize("kernel", [0, 2])
def test_d_shell_quartets(mol, basis, mode, oflags, kernel):
    """Cartesian d shells (total L up to 8, Boys order up to 8): kernel 0 = the class engine (register-
    resident slice kernels, warp-per-quartet fallback for in-core builds of L >= 6 classes), kernel 2 = the
    general shell-quartet list; tensor, G (in-core and direct) against the oracle."""
    fb = flat_for(mol, basis)
    ob = oracle.OracleBasis.from_flat(fb)
    want, _ = oracle.from_basis_packed(ob, oflags, fast=True)
    D = random_symmetric_density(fb.n_basis, seed=12)
    with FockBuilder(fb, mode=mode, kernel=kernel) as b:
        b.eri_build()
        got = b.eri_export_packed()
        G = b.build_g(D)
        st = b.stats()
    assert (st["general_shell_quartets_evaluated"] > 0) == (kernel == 2) and st["generic_integrals_evaluated"] == 0
    assert np.max(np.abs(got - want)) <= ERI_TOL
    assert_same_cut(got, want)
    G_want = oracle.build_g(want, fb.n_basis, D)
    assert np.max(np.abs(G - G_want)) <= G_TOL
    with FockBuilder(fb, mode=mode, storage=_lib.STORAGE_DIRECT, kernel=kernel) as b:
        Gd = b.eri_build().build_g(D)
    assert np.max(np.abs(Gd - G_want)) <= G_TOL


def test_scf_with_d_shells_water_631gss():
    atoms = atoms_for("water", "6-31G**")
    r = try_scf(atoms, 100, 1e-6, 0, engine=FockEngine())
    ro = try_scf(atoms, 100, 1e-6, 0, engine=OracleEngine(oracle.FAITHFUL, fast=True))
    assert r is not None and ro is not None and r.iterations == ro.iterations
    assert abs(r.total_energy - ro.total_energy) <= E_TOL


def test_device_resident_scf_matches_host_loop():
    """the SCF step with D/F/DIIS resident on the GPU (cuBLAS/cuSOLVER through torch)
    gives the same energies and iteration counts as the host loop."""
    import torch

    from fock_rs_b200.scf import try_scf_device

    for mol, basis, energy, iters in [("water", "STO-3G", -74.9874357034, 7), ("water", "6-31G", -76.0126589918, 12)]:
        atoms = atoms_for(mol, basis)
        with torch.cuda.stream(torch.cuda.Stream()):
            r = try_scf_device(atoms, 100, 1e-6, 0)
        assert r is not None and r.iterations == iters
        assert abs(r.total_energy - energy) <= 2e-8
