"""Molecular-orbital evaluation and the renderer's voxel grid:
ContractedGaussian::evaluate (basis/src/contracted_gaussian.rs:320-331), MolecularOrbital::evaluate and the
|c| > 0.05 filter (scf/src/molecular_orbitals.rs:37-43,70-79), State::create_wave_texture_data
(volumetric/src/main.rs:100-125).  CPU tests pin the oracle; GPU tests hold fock_orbital_eval /
fock_orbital_voxelize (through the C ABI) to the oracle and to tests/golden/orbital.npz.

Tolerances: FP64 values within 1e-13 absolute (|psi| <= ~10; the only divergence from the reference's
arithmetic is the last ulp of exp()); the f32 grid within one f32 ulp (+ 1e-13 next to nodal surfaces) of the oracle's,
>= 99.5 % of voxels bit-identical."""
import os

import numpy as np
import pytest

import oracle
from helpers import flat_for

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("water", "STO-3G"), ("water", "6-31G"), ("water", "6-31G**"), ("benzene", "6-31G")]


def coeff_for(n, seed=20261017):
    """A synthetic coefficient column: a spread of magnitudes either side of the 0.05 filter."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1.0, 1.0, n)
    c[rng.uniform(size=n) < 0.3] *= 0.04  # below MIN_COEFFICIENT_MAGNITUDE: dropped
    return c


def points_for(fb, n=4000, seed=7):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-10.0, 10.0, (n, 3))
    cen = np.asarray(fb.center).reshape(-1, 3)
    return np.concatenate([pts, cen[:4], cen[:4] + 1e-3])  # also on / next to nuclei


def numpy_orbital(fb, coeff, pts, min_coeff):
    """Independent vectorised evaluation (not the reference's summation order)."""
    cen = np.asarray(fb.center).reshape(-1, 3)
    ang = np.asarray(fb.angular).reshape(-1, 3)
    off = np.asarray(fb.prim_offset)
    out = np.zeros(len(pts))
    for f in range(fb.n_basis):
        if not abs(coeff[f]) > min_coeff:
            continue
        d = cen[f] - pts
        r2 = (d * d).sum(1)
        poly = d[:, 0] ** ang[f, 0] * d[:, 1] ** ang[f, 1] * d[:, 2] ** ang[f, 2]
        a = np.asarray(fb.exponent[off[f]:off[f + 1]])
        c = np.asarray(fb.coefficient[off[f]:off[f + 1]])
        out += coeff[f] * poly * (c[None, :] * np.exp(-a[None, :] * r2[:, None])).sum(1)
    return out


# ----------------------------------------------------------------------------- CPU: the oracle

@pytest.mark.parametrize("mol,basis", CASES)
def test_oracle_orbital_vs_numpy(mol, basis):
    fb = flat_for(mol, basis)
    ob = oracle.OracleBasis.from_flat(fb)
    c = coeff_for(fb.n_basis)
    pts = points_for(fb, 500)
    for mc in (0.05, -1.0):
        got = oracle.orbital_eval(ob, c, pts, mc)
        want = numpy_orbital(fb, c, pts, mc)
        assert np.max(np.abs(got - want)) <= 1e-12


def test_oracle_orbital_known_answers():
    fb = flat_for("water", "STO-3G")  # H(1s), O(1s, 2s, 2p_z, 2p_y, 2p_x), H(1s); BasisSet::get order
    ob = oracle.OracleBasis.from_flat(fb)
    cen = np.asarray(fb.center).reshape(-1, 3)
    ang = np.asarray(fb.angular).reshape(-1, 3)
    off = np.asarray(fb.prim_offset)
    # one s function at its own centre: sum of the (norm-folded) coefficients
    e = np.zeros(fb.n_basis)
    e[0] = 1.0
    v = oracle.orbital_eval(ob, e, cen[0:1], 0.05)[0]
    assert v == pytest.approx(float(np.sum(fb.coefficient[off[0]:off[1]])), rel=1e-15)
    # diff = position - at (contracted_gaussian.rs:321): a p function is NEGATIVE on its positive axis
    p = int(np.where(ang.sum(1) == 1)[0][0])
    axis = int(np.argmax(ang[p]))
    e = np.zeros(fb.n_basis)
    e[p] = 1.0
    at = cen[p].copy()
    at[axis] += 0.5
    assert oracle.orbital_eval(ob, e, at[None, :], 0.05)[0] < 0.0
    # strict '>' in MolecularOrbitals::new: |c| == 0.05 is dropped, the next double up is kept
    e = np.zeros(fb.n_basis)
    e[0] = 0.05
    assert oracle.orbital_eval(ob, e, cen[0:1], 0.05)[0] == 0.0
    e[0] = np.nextafter(0.05, 1.0)
    assert oracle.orbital_eval(ob, e, cen[0:1], 0.05)[0] != 0.0


def test_oracle_voxel_coordinates():
    """volumetric/src/main.rs:108-118: voxel i = x + n*y + n*n*z sits at ((x,y,z)/n - 0.5) * box."""
    fb = flat_for("water", "6-31G")
    ob = oracle.OracleBasis.from_flat(fb)
    c = coeff_for(fb.n_basis)
    n, box = 12, 20.0
    grid = oracle.voxelize(ob, c, n, box)
    assert grid.dtype == np.float32 and grid.shape == (n ** 3,)
    idx = np.arange(n ** 3)
    pts = np.stack([(idx % n) / n - 0.5, ((idx // n) % n) / n - 0.5, (idx // (n * n)) / n - 0.5], 1) * box
    assert np.array_equal(grid, oracle.orbital_eval(ob, c, pts).astype(np.float32))
    assert np.array_equal(oracle.voxelize(ob, c, n, box, begin=100, end=300), grid[100:300])


def test_oracle_orbital_golden():
    g = np.load(os.path.join(GOLD, "orbital.npz"))
    for mol, basis in CASES:
        tag = f"{mol}_{basis}".replace("*", "s").replace("-", "")
        fb = flat_for(mol, basis)
        ob = oracle.OracleBasis.from_flat(fb)
        assert np.array_equal(oracle.orbital_eval(ob, g[f"coeff_{tag}"], g["points"]), g[f"values_{tag}"])
    fb = flat_for("water", "6-31G")
    assert np.array_equal(oracle.voxelize(oracle.OracleBasis.from_flat(fb), g["coeff_water_631G"], 16, 20.0), g["grid16_water_631G"])


# ----------------------------------------------------------------------------- GPU: through the C ABI

@pytest.mark.gpu
@pytest.mark.parametrize("mol,basis", CASES)
def test_cuda_orbital_eval_vs_oracle(mol, basis):
    from fock_rs_b200.fock import FockBuilder

    fb = flat_for(mol, basis)
    ob = oracle.OracleBasis.from_flat(fb)
    c = coeff_for(fb.n_basis)
    pts = points_for(fb)
    with FockBuilder(fb) as b:
        for mc in (0.05, -1.0, 10.0):
            got = b.orbital_eval(c, pts, mc)
            want = oracle.orbital_eval(ob, c, pts, mc)
            assert np.max(np.abs(got - want)) <= 1e-13
        assert np.all(b.orbital_eval(c, pts, 10.0) == 0.0)  # every function filtered out
        assert b.orbital_eval(c, np.zeros((0, 3))).shape == (0,)  # empty batch


@pytest.mark.gpu
def test_cuda_orbital_matches_golden():
    from fock_rs_b200.fock import FockBuilder

    g = np.load(os.path.join(GOLD, "orbital.npz"))
    for mol, basis in CASES:
        tag = f"{mol}_{basis}".replace("*", "s").replace("-", "")
        with FockBuilder(flat_for(mol, basis)) as b:
            got = b.orbital_eval(g[f"coeff_{tag}"], g["points"])
            assert np.max(np.abs(got - g[f"values_{tag}"])) <= 1e-13
            if tag == "water_631G":
                grid = b.orbital_voxelize(g[f"coeff_{tag}"], 16, 20.0)
                _assert_f32_grid_close(grid, g["grid16_water_631G"])


def _assert_f32_grid_close(got, want):
    assert got.dtype == np.float32 and got.shape == want.shape
    same = got.view(np.uint32) == want.view(np.uint32)
    # where not bit-identical: adjacent f32 values (the FP64 psi differed in the last ulp of exp()), or -- next to
    # a nodal surface, where the terms cancel -- within the FP64 tolerance of 1e-13 absolute
    bad = ~same
    if bad.any():
        assert np.all(np.abs(got[bad].astype(np.float64) - want[bad]) <= np.spacing(np.abs(want[bad])).astype(np.float64) * 1.0001 + 1e-13)
    assert same.mean() >= 0.995


@pytest.mark.gpu
def test_cuda_voxelize_reference_grid():
    """The renderer's own size: 200^3 voxels, box 20 (volumetric/src/main.rs:63-64), benzene/6-31G.
    The whole grid is produced on the GPU; the oracle checks four z-slabs (bounded CPU time) and a
    size-independent property covers the rest: psi is linear in the coefficient column."""
    from fock_rs_b200.fock import FockBuilder

    fb = flat_for("benzene", "6-31G")
    ob = oracle.OracleBasis.from_flat(fb)
    c = coeff_for(fb.n_basis)
    n, box = 200, 20.0
    with FockBuilder(fb) as b:
        grid = b.orbital_voxelize(c, n, box)
        st = b.orbital_stats()
        assert st["functions"] == int(np.sum(np.abs(c) > 0.05)) and 0 < st["groups"] <= st["functions"]
        for z in (0, 77, 100, 199):
            lo, hi = z * n * n, (z + 1) * n * n
            _assert_f32_grid_close(grid[lo:hi], oracle.voxelize(ob, c, n, box, begin=lo, end=hi))
        assert np.isfinite(grid).all() and np.abs(grid).max() > 1e-3
        # linearity on the FP64 path at the voxel centres of one x-row through the molecule
        idx = np.arange(n) + n * 100 + n * n * 100
        pts = np.stack([(idx % n) / n - 0.5, ((idx // n) % n) / n - 0.5, (idx // (n * n)) / n - 0.5], 1) * box
        c2 = coeff_for(fb.n_basis, seed=3)
        lhs = b.orbital_eval(c + c2, pts, -1.0)
        rhs = b.orbital_eval(c, pts, -1.0) + b.orbital_eval(c2, pts, -1.0)
        assert np.max(np.abs(lhs - rhs)) <= 1e-12
        assert np.array_equal(grid[idx], b.orbital_eval(c, pts).astype(np.float32))
        assert b.orbital_voxelize(c, 0, box).shape == (0,)


@pytest.mark.gpu
def test_cuda_orbital_bad_args():
    from fock_rs_b200 import _lib
    from fock_rs_b200.fock import FockBuilder

    fb = flat_for("water", "STO-3G")
    with FockBuilder(fb) as b:
        with pytest.raises(ValueError):
            b.orbital_eval(np.zeros(fb.n_basis + 1), np.zeros((1, 3)))
        with pytest.raises(_lib.FockError):
            b.orbital_voxelize(np.ones(fb.n_basis), -1, 20.0)


@pytest.mark.gpu
def test_molecular_orbitals_mirror_after_scf():
    """orbitals[n].evaluate(at) for the converged H2O/STO-3G coefficient matrix, as pointcloud/src/main.rs:108 calls it."""
    from helpers import atoms_for
    from fock_rs_b200.fock import FockEngine, MolecularOrbitals
    from fock_rs_b200.scf import try_scf

    atoms = atoms_for("water", "STO-3G")
    res = try_scf(list(atoms), 100, 1e-6, 0, engine=FockEngine())
    assert res is not None
    fb = flat_for("water", "STO-3G")
    ob = oracle.OracleBasis.from_flat(fb)
    mos = MolecularOrbitals(fb, res.coefficient_matrix)
    assert len(mos) == fb.n_basis
    rng = np.random.default_rng(1)
    pts = rng.uniform(-3.0, 3.0, (64, 3))
    for n in (0, 4, fb.n_basis - 1):
        col = np.ascontiguousarray(res.coefficient_matrix[:, n])
        want = oracle.orbital_eval(ob, col, pts)
        assert np.max(np.abs(mos[n].evaluate_many(pts) - want)) <= 1e-13
        assert abs(mos[n].evaluate(pts[0]) - want[0]) <= 1e-13
        _assert_f32_grid_close(mos[n].voxelize(24, 12.0), oracle.voxelize(ob, col, 24, 12.0))


@pytest.mark.gpu
def test_cuda_orbital_deep_contraction_path():
    """Contractions longer than the kernel's 8-entry exponential cache (the 9-primitive s functions of the cc-pVDZ-shaped
    set) take the per-function branch, which recomputes exp() per function exactly as the reference does."""
    from fock_rs_b200.fock import FockBuilder

    fb = flat_for("water", "cc-pVDZ")
    assert int(np.max(np.diff(np.asarray(fb.prim_offset)))) == 9
    ob = oracle.OracleBasis.from_flat(fb)
    c = coeff_for(fb.n_basis, seed=12)
    c[5] = 0.8  # make sure both 9-primitive functions are kept
    c[6] = -0.6
    pts = points_for(fb, 2000)
    with FockBuilder(fb) as b:
        got = b.orbital_eval(c, pts)
        grid = b.orbital_voxelize(c, 32, 8.0)
    want = oracle.orbital_eval(ob, c, pts)
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, float(np.max(np.abs(want))))
    _assert_f32_grid_close(grid, oracle.voxelize(ob, c, 32, 8.0))
