"""The header-only C++ mirror (fock_rs_b200/hostcpp/fock_b200.hpp) compiles against the C ABI and,
on a GPU box, reproduces the oracle for H2/STO-3G (the molecule of the reference's test_integrals)."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from fock_rs_b200 import _lib
from fock_rs_b200.basis_set import BasisSet, flatten

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "hostcpp_main")


def _compile():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "hostcpp_main.cpp"),
                           "-o", EXE, "-pthread", "-L" + lib_dir, "-lfock_b200", "-Wl,-rpath," + lib_dir, "-Wl,-rpath,/usr/local/cuda/lib64"])


def test_hostcpp_compiles_and_links():
    _lib.load()
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_hostcpp_matches_oracle():
    _compile()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    eris = [float(x) for x in lines[0].split()]
    g = np.array([float(x) for x in lines[1].split()]).reshape(2, 2)
    bs = BasisSet.named("STO-3G")
    fb = flatten(bs.get((0.0, 0.0, 0.0), 1).basis + bs.get((0.0, 0.0, 1.2), 1).basis)
    ob = oracle.OracleBasis.from_flat(fb)
    want = [oracle.eri(ob, 0, 0, 0, 0), oracle.eri(ob, 1, 0, 1, 0), oracle.eri(ob, 1, 1, 0, 0)]
    np.testing.assert_allclose(eris, want, rtol=0, atol=1e-12)
    packed, _ = oracle.from_basis_packed(ob, oracle.FAITHFUL, fast=True)
    np.testing.assert_allclose(g, oracle.build_g(packed, 2, np.eye(2)), rtol=0, atol=1e-12)
    psi = [float(x) for x in lines[2].split()]
    c = np.array([0.55, 0.55])
    want_psi = oracle.orbital_eval(ob, c, np.array([[0.0, 0.0, 0.6], [0.3, -0.2, 0.1]]))
    np.testing.assert_allclose(psi[:2], want_psi, rtol=0, atol=1e-13)
    grid = oracle.voxelize(ob, c, 2, 4.0)
    np.testing.assert_allclose(psi[2:], [grid[0], grid[7]], rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_hostcpp_two_gpus_complete_g_through_the_abi():
    """The collective behind the C ABI, driven from compiled host code only: 2 threads / 2 GPUs / one NCCL
    communicator attached with fock_comm_init; every rank's fock_build_g returns the complete G."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _compile()
    out = subprocess.run([EXE, "multi", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    tag, d1, dr = out.stdout.strip().splitlines()[-1].split()
    assert tag == "MULTI_OK"
    assert float(d1) <= 1e-11  # same integrals, different summation order
    assert float(dr) == 0.0    # every rank holds the same all-reduced matrix
