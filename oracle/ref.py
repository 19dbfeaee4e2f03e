"""ctypes binding of oracle/fock_oracle.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfock_oracle.so")

BOYS_EXACT = 1
FIX_R = 2
NO_CUT = 4
FAITHFUL = 0
ACCURATE = BOYS_EXACT | FIX_R | NO_CUT


def build(force=False):
    src = os.path.join(_HERE, "fock_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class _Basis(C.Structure):
    _fields_ = [
        ("n_basis", C.c_int32),
        ("center", C.POINTER(C.c_double)),
        ("angular", C.POINTER(C.c_int32)),
        ("prim_offset", C.POINTER(C.c_int32)),
        ("exponent", C.POINTER(C.c_double)),
        ("coefficient", C.POINTER(C.c_double)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        bp = C.POINTER(_Basis)
        L.oracle_boys.restype = C.c_double
        L.oracle_boys.argtypes = [C.c_double, C.c_int]
        L.oracle_boys_exact.restype = C.c_double
        L.oracle_boys_exact.argtypes = [C.c_double, C.c_int]
        L.oracle_hermite_expansion.restype = C.c_double
        L.oracle_hermite_expansion.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.oracle_coulomb_auxiliary.restype = C.c_double
        L.oracle_coulomb_auxiliary.argtypes = [C.c_int] * 4 + [C.c_double, dp, C.c_double, C.c_int]
        L.oracle_norm.restype = C.c_double
        L.oracle_norm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
        L.oracle_eri.restype = C.c_double
        L.oracle_eri.argtypes = [bp] + [C.c_int] * 5
        L.oracle_eri_fast.restype = C.c_double
        L.oracle_eri_fast.argtypes = [bp] + [C.c_int] * 5
        L.oracle_packed_index.restype = C.c_size_t
        L.oracle_packed_index.argtypes = [C.c_size_t] * 4
        L.oracle_packed_len.restype = C.c_size_t
        L.oracle_packed_len.argtypes = [C.c_size_t]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_from_basis_packed.restype = C.c_int64
        L.oracle_from_basis_packed.argtypes = [bp, C.c_int, C.c_int, dp]
        L.oracle_sample_rows.restype = C.c_int64
        L.oracle_sample_rows.argtypes = [bp, C.c_int, C.c_int, dp, C.c_int64, C.c_int64, C.c_int64, dp]
        L.oracle_diagonal.restype = None
        L.oracle_diagonal.argtypes = [bp, C.c_int, C.c_int, dp]
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_direct_g.restype = C.c_int64
        L.oracle_direct_g.argtypes = [bp, C.c_int, C.c_int, dp, dp, dp, C.c_int64, C.c_int64, C.c_int64]
        L.oracle_build_g.restype = None
        L.oracle_build_g.argtypes = [dp, C.c_int, dp, dp, C.c_int]
        L.oracle_one_electron.restype = None
        L.oracle_one_electron.argtypes = [bp, C.c_int, dp, dp, C.c_int, dp, dp, dp]
        L.oracle_orbital_eval.argtypes = [bp, dp, C.c_double, dp, C.c_int64, dp]
        L.oracle_orbital_eval.restype = None
        L.oracle_voxelize.argtypes = [bp, dp, C.c_double, C.c_int, C.c_double, C.c_int64, C.c_int64, C.POINTER(C.c_float)]
        L.oracle_voxelize.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleBasis:
    """Holds the flat SoA (same arrays as fock_basis in include/fock_b200.h)."""

    def __init__(self, center, angular, prim_offset, exponent, coefficient):
        self.center = np.ascontiguousarray(center, dtype=np.float64).reshape(-1, 3)
        self.angular = np.ascontiguousarray(angular, dtype=np.int32).reshape(-1, 3)
        self.prim_offset = np.ascontiguousarray(prim_offset, dtype=np.int32)
        self.exponent = np.ascontiguousarray(exponent, dtype=np.float64)
        self.coefficient = np.ascontiguousarray(coefficient, dtype=np.float64)
        self.n = self.center.shape[0]
        self.c = _Basis(
            self.n,
            _dp(self.center),
            self.angular.ctypes.data_as(C.POINTER(C.c_int32)),
            self.prim_offset.ctypes.data_as(C.POINTER(C.c_int32)),
            _dp(self.exponent),
            _dp(self.coefficient),
        )

    @classmethod
    def from_flat(cls, flat):
        """`flat` = fock_rs_b200.basis_set.FlatBasis (or anything with the same five arrays)."""
        return cls(flat.center, flat.angular, flat.prim_offset, flat.exponent, flat.coefficient)


def boys(t, n):
    return lib().oracle_boys(float(t), int(n))


def boys_exact(t, n):
    return lib().oracle_boys_exact(float(t), int(n))


def hermite_expansion(ijt, dist, a, b):
    i, j, t = ijt
    return lib().oracle_hermite_expansion(i, j, t, float(dist), float(a), float(b))


def coulomb_auxiliary(t, u, v, n, p, diff, dist_sq, flags=0):
    d = np.ascontiguousarray(diff, dtype=np.float64)
    return lib().oracle_coulomb_auxiliary(t, u, v, n, float(p), _dp(d), float(dist_sq), flags)


def norm(ang, alpha):
    return lib().oracle_norm(int(ang[0]), int(ang[1]), int(ang[2]), float(alpha))


def eri(ob, a, b, c, d, flags=0, fast=False):
    f = lib().oracle_eri_fast if fast else lib().oracle_eri
    return f(C.byref(ob.c), a, b, c, d, flags)


def packed_index(x, y, z, w):
    return lib().oracle_packed_index(x, y, z, w)


def packed_len(n):
    return lib().oracle_packed_len(n)


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    """OpenMP threads of the oracle's parallel loops (torchrun exports OMP_NUM_THREADS=1)."""
    lib().oracle_set_num_threads(int(n))


def direct_g(ob, D, flags=0, fast=True, diag=None, row_start=0, row_stride=1, max_rows=None):
    """Tensor build + G loop in integral-direct form (no N^4 storage); the whole job by default, or the
    canonical rows row_start, row_start+row_stride, ...  Returns (G, n_evaluated).  D must be symmetric."""
    n = ob.n
    Dc = np.asfortranarray(D, dtype=np.float64)
    assert np.array_equal(Dc, Dc.T), "direct_g needs a symmetric density"
    if diag is None:
        diag = diagonal(ob, flags, fast)
    d = np.ascontiguousarray(diag, dtype=np.float64)
    G = np.zeros((n, n), dtype=np.float64, order="F")
    npair = n * (n + 1) // 2
    nev = lib().oracle_direct_g(C.byref(ob.c), flags, int(fast), _dp(d), _dp(Dc), _dp(G), int(row_start), int(row_stride),
                                int(npair if max_rows is None else max_rows))
    return np.array(G), int(nev)


def from_basis_packed(ob, flags=0, fast=True):
    """ElectronTensor::from_basis in canonical packed form.  Returns (packed, n_evaluated)."""
    out = np.zeros(packed_len(ob.n), dtype=np.float64)
    nev = lib().oracle_from_basis_packed(C.byref(ob.c), flags, int(fast), _dp(out))
    return out, int(nev)


def diagonal(ob, flags=0, fast=True):
    out = np.zeros(ob.n * (ob.n + 1) // 2, dtype=np.float64)
    lib().oracle_diagonal(C.byref(ob.c), flags, int(fast), _dp(out))
    return out


def sample_rows(ob, diag, row_start, row_stride, max_rows, flags=0, fast=False):
    cs = C.c_double(0.0)
    d = np.ascontiguousarray(diag, dtype=np.float64)
    nev = lib().oracle_sample_rows(C.byref(ob.c), flags, int(fast), _dp(d), row_start, row_stride, max_rows, C.byref(cs))
    return int(nev), cs.value


def build_g(packed, n, D, parallel=True):
    """G (N x N, symmetric) from the packed tensor and a density, reference summation order."""
    Dc = np.asfortranarray(D, dtype=np.float64)
    G = np.zeros((n, n), dtype=np.float64, order="F")
    p = np.ascontiguousarray(packed, dtype=np.float64)
    lib().oracle_build_g(_dp(p), n, _dp(Dc), _dp(G), int(parallel))
    return np.array(G)


def one_electron(ob, nuc_pos, nuc_charge, flags=0):
    n = ob.n
    pos = np.ascontiguousarray(nuc_pos, dtype=np.float64).reshape(-1, 3)
    chg = np.ascontiguousarray(nuc_charge, dtype=np.float64)
    S = np.zeros((n, n), order="F")
    T = np.zeros((n, n), order="F")
    V = np.zeros((n, n), order="F")
    lib().oracle_one_electron(C.byref(ob.c), pos.shape[0], _dp(pos), _dp(chg), flags, _dp(S), _dp(T), _dp(V))
    return np.array(S), np.array(T), np.array(V)


def orbital_eval(ob, coeff, points, min_coeff=0.05):
    """MolecularOrbital::evaluate at a batch of points (FP64)."""
    c = np.ascontiguousarray(coeff, dtype=np.float64)
    assert c.shape == (ob.n,)
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    out = np.zeros(pts.shape[0], dtype=np.float64)
    lib().oracle_orbital_eval(C.byref(ob.c), _dp(c), float(min_coeff), _dp(pts), pts.shape[0], _dp(out))
    return out


def voxelize(ob, coeff, n_voxels=200, box_size=20.0, min_coeff=0.05, begin=0, end=None):
    """State::create_wave_texture_data for voxels [begin, end) (f32, x fastest)."""
    c = np.ascontiguousarray(coeff, dtype=np.float64)
    assert c.shape == (ob.n,)
    end = n_voxels ** 3 if end is None else end
    out = np.zeros(max(end - begin, 0), dtype=np.float32)
    lib().oracle_voxelize(C.byref(ob.c), _dp(c), float(min_coeff), int(n_voxels), float(box_size), int(begin), int(end),
                          out.ctypes.data_as(C.POINTER(C.c_float)))
    return out
