"""Test-side SCF engine backed by the CPU oracle (tests only; see oracle/fock_oracle.c)."""
import numpy as np

import oracle


class OracleEngine:
    def __init__(self, flags=oracle.FAITHFUL, fast=True):
        self.flags = flags
        self.fast = fast
        self.packed = None
        self.n = 0
        self.n_evaluated = 0

    def one_electron(self, flat, nuc_pos, nuc_chg):
        ob = oracle.OracleBasis.from_flat(flat)
        return oracle.one_electron(ob, nuc_pos, nuc_chg, self.flags)

    def from_basis(self, flat):
        ob = oracle.OracleBasis.from_flat(flat)
        self.n = ob.n
        self.packed, self.n_evaluated = oracle.from_basis_packed(ob, self.flags, self.fast)

    def build_g(self, D):
        return oracle.build_g(self.packed, self.n, np.asarray(D))
