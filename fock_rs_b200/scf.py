"""Host mirror of the reference's RHF driver `SelfConsistentField::try_scf`
(reference: scf/src/lib.rs:73-258, scf/src/utils.rs:25-138).

Everything here except the two-electron work "stays as in the reference" (north_star): the
eigensolves, DIIS and the density update run on the host with numpy/LAPACK (the reference uses
nalgebra / faer-evd).  The hot path -- `ElectronTensor::from_basis` (scf/src/lib.rs:132), the
`electron_terms` precompute (:167-173) and the per-iteration G build (:177-185) -- is delegated to an
*engine* object; the product engine is `fock_rs_b200.fock.FockEngine` (CUDA through the C ABI, no
CPU fallback).  tests/ plug the CPU oracle in through the same three-method protocol:

    engine.one_electron(flat_basis, nuc_pos, nuc_charge) -> (S, T, V)     # scf/src/lib.rs:118-128
    engine.from_basis(flat_basis)                                           # :132  (+ :167-173)
    engine.build_g(D) -> G                                                  # :177-185
"""
from __future__ import annotations

import logging
import math
import time
from collections import deque
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from .basis_set import Atom, flatten, molecule_basis

log = logging.getLogger("fock_rs_b200.scf")


@dataclass
class HartreeFockResult:
    """scf/src/lib.rs:27-37 (the MolecularOrbitals renderer helper is out of scope)."""

    orbital_energies: np.ndarray
    density_matrix: np.ndarray
    coefficient_matrix: np.ndarray
    electronic_energy: float
    total_energy: float
    iterations: int
    n_electrons: int
    n_basis: int
    energies_per_iteration: list
    errors_per_iteration: list


def hermitian(n: int, func) -> np.ndarray:
    """scf/src/utils.rs:25-28: evaluate i<=j, mirror."""
    m = np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            m[i, j] = func(i, j)
            m[j, i] = m[i, j]
    return m


def _mirror_upper(m: np.ndarray) -> np.ndarray:
    u = np.triu(m)
    return u + np.triu(m, 1).T


def sorted_eigs(m: np.ndarray):
    """scf/src/utils.rs:59-99 (faer-evd + sort ascending) -> LAPACK eigh (already ascending)."""
    w, v = np.linalg.eigh(m)
    return v, w


def diis(errors: Sequence[np.ndarray], focks: Sequence[np.ndarray]) -> Optional[np.ndarray]:
    """scf/src/utils.rs:102-138."""
    n = len(errors)
    b = np.zeros((n + 1, n + 1))
    for i in range(n):
        for j in range(n):
            b[i, j] = float(np.vdot(errors[i], errors[j]))
    b[n, :n] = -1.0
    b[:n, n] = -1.0
    rhs = np.zeros(n + 1)
    rhs[n] = -1.0
    try:
        c = np.linalg.inv(b) @ rhs
    except np.linalg.LinAlgError:
        return None
    if not np.all(np.isfinite(c)):
        return None
    out = np.zeros_like(focks[0])
    for i in range(n):
        out = out + c[i] * focks[i]
    return out


def nuclear_repulsion(atoms: Sequence[Atom]) -> float:
    """scf/src/lib.rs:98-116."""
    e = 0.0
    for i in range(len(atoms)):
        for j in range(i + 1, len(atoms)):
            d = math.dist(atoms[i].position, atoms[j].position)
            e += atoms[i].atom_type * atoms[j].atom_type / d
    return e


def try_scf(atoms: Sequence[Atom], max_iters: int, epsilon: float, molecule_charge: int = 0, engine=None) -> Optional[HartreeFockResult]:
    """`atoms.try_scf(max_iters, epsilon, molecule_charge)` (scf/src/lib.rs:73-258)."""
    if engine is None:
        from .fock import FockEngine  # CUDA; raises if the extension is missing

        engine = FockEngine()
    atoms = list(atoms)
    basis = molecule_basis(atoms)
    flat = flatten(basis)
    n_basis = flat.n_basis
    n_electrons = sum(a.electron_count() for a in atoms) - molecule_charge
    log.info("Starting SCF iteration with %d basis functions", n_basis)
    if n_electrons % 2 != 0:
        log.warning("restricted hartree fock only works properly if all orbitals are fully occupied")
    e_nuc = nuclear_repulsion(atoms)
    nuc_pos = np.array([a.position for a in atoms], dtype=np.float64)
    nuc_chg = np.array([float(a.atom_type) for a in atoms], dtype=np.float64)

    overlap, kinetic, nuclear = engine.one_electron(flat, nuc_pos, nuc_chg)

    t0 = time.perf_counter()
    engine.from_basis(flat)
    log.info("Multi-Electron tensor formation took %.1f ms", 1e3 * (time.perf_counter() - t0))

    hcore = kinetic + nuclear
    _, u = np.linalg.eigh(overlap)  # utils::eigs (nalgebra SymmetricEigen), :139-146
    diag = u.T @ (overlap @ u)
    transform = u @ (np.diag(1.0 / np.sqrt(np.diag(diag))) @ u.T)

    n_occ = n_electrons // 2

    def density_from(coeffs):
        c = coeffs[:, :n_occ]
        return _mirror_upper(2.0 * (c @ c.T))

    # extended-Hueckel guess (:150-162)
    hd = np.diag(hcore)
    h_eht = _mirror_upper(0.875 * overlap * (hd[:, None] + hd[None, :]))
    cp, _ = sorted_eigs(transform.T @ (h_eht @ transform))
    density = density_from(transform @ cp)

    prev_f, prev_e = deque(), deque()
    energies, errors = [], []
    t0 = time.perf_counter()
    for it in range(max_iters + 1):
        guess = engine.build_g(density)  # :177-185
        fock = hcore + guess
        err_m = fock @ density @ overlap - overlap @ density @ fock
        error = float(np.max(np.abs(err_m)))
        prev_e.append(err_m)
        prev_f.append(fock)
        if len(prev_e) > 16:
            prev_e.popleft()
            prev_f.popleft()
        if len(prev_f) < 5:
            fock_x = prev_f[-1].copy()
        else:
            fock_x = diis(prev_e, prev_f)
            if fock_x is None:
                fock_x = prev_f[-1].copy()
        cp, orb_e = sorted_eigs(transform.T @ (fock_x @ transform))
        coeffs = transform @ cp
        density = density_from(coeffs)
        e_el = 0.5 * float(np.trace(density @ (2.0 * hcore + guess)))
        energies.append(e_el + e_nuc)
        errors.append(error)
        if error < epsilon or it == max_iters:
            log.info("SCF took %.4f s to converge (%d)", time.perf_counter() - t0, it)
            return HartreeFockResult(orb_e, density, coeffs, e_el, e_nuc + e_el, it, n_electrons, n_basis, energies, errors)
        elif not (math.isfinite(error) and error >= 2.2250738585072014e-308):  # !is_normal()
            return None
        log.debug("Iteration %d: max error (loss): %.5e, energy: %.5f", it, error, e_el + e_nuc)
    return None


def try_scf_device(atoms: Sequence[Atom], max_iters: int, epsilon: float, molecule_charge: int = 0, mode=0, screen=0, storage=0,
                   device: int = 0) -> Optional[HartreeFockResult]:
    """Device-resident variant of `try_scf`: the density, Fock and DIIS matrices stay
    in HBM, FDS - SDF / the similarity transforms run through cuBLAS and the two eigensolves through
    cuSOLVER (torch.linalg), and G comes from `fock_build_g_device` -- no N x N matrix crosses PCIe inside
    the loop; one scalar (the DIIS error) is read back per iteration for the convergence test.
    Same algorithm and constants as `try_scf` (scf/src/lib.rs:73-258)."""
    import torch

    from .fock import FockBuilder

    atoms = list(atoms)
    flat = flatten(molecule_basis(atoms))
    n = flat.n_basis
    n_electrons = sum(a.electron_count() for a in atoms) - molecule_charge
    n_occ = n_electrons // 2
    e_nuc = nuclear_repulsion(atoms)
    dev = torch.device("cuda", device)
    builder = FockBuilder(flat, mode=mode, screen=screen, storage=storage, device=device, keep_eri=False)
    try:
        stream = torch.cuda.current_stream(dev)
        own_stream = stream.cuda_stream == 0  # legacy default stream: the library keeps its own and we fence around it
        if not own_stream:
            builder.set_stream(stream.cuda_stream)
        nuc_pos = np.array([a.position for a in atoms], dtype=np.float64)
        nuc_chg = np.array([float(a.atom_type) for a in atoms], dtype=np.float64)
        S_h, T_h, V_h = builder.one_electron(nuc_pos, nuc_chg)
        builder.eri_build()
        S = torch.from_numpy(np.ascontiguousarray(S_h)).to(dev)
        H = torch.from_numpy(np.ascontiguousarray(T_h + V_h)).to(dev)

        def mirror_upper(m):
            u = torch.triu(m)
            return u + torch.triu(m, 1).T

        _, u = torch.linalg.eigh(S)
        diag = u.T @ (S @ u)
        X = u @ (torch.diag(1.0 / torch.sqrt(torch.diagonal(diag))) @ u.T)

        def density_from(coeffs):
            c = coeffs[:, :n_occ]
            return mirror_upper(2.0 * (c @ c.T))

        hd = torch.diagonal(H)
        h_eht = mirror_upper(0.875 * S * (hd[:, None] + hd[None, :]))
        _, cp = torch.linalg.eigh(X.T @ (h_eht @ X))
        D = density_from(X @ cp).contiguous()
        G = torch.empty_like(D)
        prev_f, prev_e = deque(), deque()
        energies, errors = [], []
        for it in range(max_iters + 1):
            if own_stream:
                torch.cuda.synchronize(dev)
            builder.build_g_device(D.data_ptr(), G.data_ptr())  # symmetric: row- and column-major images coincide
            if own_stream:
                torch.cuda.synchronize(dev)
            guess = G.clone()
            fock = H + guess
            err_m = fock @ D @ S - S @ D @ fock
            error = float(torch.max(torch.abs(err_m)))  # the one scalar read back per iteration
            prev_e.append(err_m)
            prev_f.append(fock)
            if len(prev_e) > 16:
                prev_e.popleft()
                prev_f.popleft()
            fock_x = fock
            if len(prev_f) >= 5:
                m = len(prev_e)
                E = torch.stack([e.reshape(-1) for e in prev_e])
                B = torch.zeros((m + 1, m + 1), dtype=torch.float64, device=dev)
                B[:m, :m] = E @ E.T
                B[m, :m] = -1.0
                B[:m, m] = -1.0
                rhs = torch.zeros(m + 1, dtype=torch.float64, device=dev)
                rhs[m] = -1.0
                try:
                    c = torch.linalg.inv(B) @ rhs
                    if bool(torch.all(torch.isfinite(c))):
                        fock_x = sum(c[i] * prev_f[i] for i in range(m))
                except Exception:
                    pass
            orb_e, cp = torch.linalg.eigh(X.T @ (fock_x @ X))
            coeffs = X @ cp
            D = density_from(coeffs).contiguous()
            e_el = 0.5 * float(torch.trace(D @ (2.0 * H + guess)))
            energies.append(e_el + e_nuc)
            errors.append(error)
            if error < epsilon or it == max_iters:
                return HartreeFockResult(orb_e.cpu().numpy(), D.cpu().numpy(), coeffs.cpu().numpy(), e_el, e_nuc + e_el, it, n_electrons, n,
                                         energies, errors)
            elif not (math.isfinite(error) and error >= 2.2250738585072014e-308):
                return None
        return None
    finally:
        builder.close()
