"""CPU-side tests: host mirror of the reference's data types, the C-ABI surface (symbols only -- no
compute without a GPU), and the multi-process plumbing on gloo."""
import os
import re

import numpy as np
import pytest

from fock_rs_b200 import _lib, molecules, parallel
from fock_rs_b200.basis_set import BasisSet, GaussianPrimitive, flatten, molecule_basis
from fock_rs_b200.scf import diis, hermitian
from fock_rs_b200.xyz import parse_xyz
from helpers import atoms_for, flat_for

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "fock_b200.h")).read()
    declared = set(re.findall(r"\b(fock_[a-z0-9_]+)\s*\(", header))
    declared -= {"fock_ctx"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fock_abi_version() == 5
    assert lib.fock_packed_len(7) == 406 and lib.fock_packed_len(66) == 2445366


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fock_rs_b200.fock import FockBuilder

    with pytest.raises(_lib.FockError) as ei:
        FockBuilder(flat_for("water", "STO-3G"))
    assert ei.value.code == _lib.ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fock_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "fock_oracle" not in src, f


def test_basis_set_get_matches_reference_conventions():
    # all shells Cartesian, components x-outer..z-inner: p = (z, y, x), d = (zz, yz, yy, xz, xy, xx)
    atom = BasisSet.named("6-31G**").get((0.0, 0.0, 0.0), 8)
    angs = [f.primitives[0].angular for f in atom.basis]
    assert len(atom.basis) == 15
    assert angs[1:4] == [(0, 0, 0), (0, 0, 1), (0, 1, 0)]  # SP shell: s then p_z, p_y, ...
    assert angs[-6:] == [(0, 0, 2), (0, 1, 1), (0, 2, 0), (1, 0, 1), (1, 1, 0), (2, 0, 0)]
    # primitives individually normalised (basis/src/primitives.rs:19-27)
    a = 0.8
    assert GaussianPrimitive.norm((0, 0, 0), a) == pytest.approx((2 * a / np.pi) ** 0.75, rel=1e-15)
    assert GaussianPrimitive.norm((1, 1, 0), a) / GaussianPrimitive.norm((2, 0, 0), a) == pytest.approx(np.sqrt(3.0), rel=1e-15)


def test_xyz_reader_uses_reference_conversion_and_symbols():
    atoms = parse_xyz("2\ncomment line\nH 0 0 1.0\nO 0.5 -0.25 0\n", BasisSet.named("STO-3G"))
    assert [a.atom_type for a in atoms] == [1, 8]
    assert atoms[0].position == (0.0, 0.0, 1.89)
    assert atoms[1].position == (0.5 * 1.89, -0.25 * 1.89, 0.0)
    with pytest.raises(NotImplementedError):
        parse_xyz("Cl 0 0 0\n", BasisSet.named("STO-3G"))


def test_workload_sizes_match_baseline_configs():
    assert flat_for("water", "STO-3G").n_basis == 7
    assert flat_for("benzene", "6-31G").n_basis == 66
    assert flat_for("w16", "6-31G").n_basis == 208
    assert flat_for("alkane40", "6-31G").n_basis == 524
    assert flat_for("w32", "6-31G**").n_basis == 800
    assert len(atoms_for("alkane40", "6-31G")) == 122


def test_flatten_layout():
    fb = flat_for("water", "6-31G")
    assert fb.prim_offset[0] == 0 and fb.prim_offset[-1] == fb.exponent.size == fb.coefficient.size
    assert fb.center.shape == (13, 3) and fb.angular.shape == (13, 3)
    # functions of one SP shell share exponents
    np.testing.assert_array_equal(fb.exponent[fb.prim_offset[3] : fb.prim_offset[4]], fb.exponent[fb.prim_offset[4] : fb.prim_offset[5]])


def test_hermitian_and_diis_host_logic():
    m = hermitian(3, lambda i, j: i * 10 + j)
    assert m[2, 0] == m[0, 2] == 2 and m[1, 2] == 12
    # DIIS with two error vectors e and -e/2 extrapolates to the zero-error combination
    f0, f1 = np.eye(2), 3 * np.eye(2)
    e0, e1 = np.array([[0.0, 1.0], [-1.0, 0.0]]), np.array([[0.0, -0.5], [0.5, 0.0]])
    out = diis([e0, e1], [f0, f1])
    np.testing.assert_allclose(out, (1 / 3) * f0 + (2 / 3) * f1, atol=1e-12)
    assert diis([e0, e0], [f0, f1]) is None or np.all(np.isfinite(diis([e0, e0], [f0, f1])))


def test_digest_ranges_partition():
    n = 2362314
    for world in (1, 2, 3, 8):
        r = parallel.digest_ranges(n, world)
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert max(e - b for b, e in r) - min(e - b for b, e in r if e > b) <= world


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fb = flat_for("water", "STO-3G")
        ob = oracle.OracleBasis.from_flat(fb)
        packed, _ = oracle.from_basis_packed(ob, oracle.FAITHFUL, fast=True)
        rng = np.random.default_rng(1)
        D = rng.uniform(-1, 1, (7, 7))
        D = 0.5 * (D + D.T)
        full = oracle.build_g(packed, 7, D)
        # the path shards over independent integrals: rank r keeps packed slots e % world == r, its
        # partial G is linear in them, and one all-reduce of N x N doubles restores the full G
        mine = np.where(np.arange(packed.size) % world == rank, packed, 0.0)
        partial = oracle.build_g(mine, 7, D)
        total = parallel.allreduce_host_matrix(partial)
        q.put((rank, float(np.max(np.abs(total - full)))))
    finally:
        dist.destroy_process_group()


def test_partial_g_allreduce_on_gloo_world2():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(err <= 1e-12 for _, err in res)


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` needs no GPU: exactly one line on stdout, JSON, with the driver's keys; anything else
    the run prints goes to stderr."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "water", "--steps", "1",
                          "--warmup", "0", "--ref-budget", "0.2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_bench_reference_arm_other_ranks_exit_quietly():
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "water", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
