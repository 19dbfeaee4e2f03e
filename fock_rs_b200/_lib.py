"""ctypes loader of libfock_b200.so (the C ABI in include/fock_b200.h).

There is no CPU fallback: if the CUDA library is not built, importing the compute layer fails
loudly (build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C fock_rs_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FOCK_B200_LIB: measure an experimental build of the same ABI (tools/); the shipped library is the in-tree one
LIB_PATH = os.environ.get("FOCK_B200_LIB") or os.path.join(_HERE, "libfock_b200.so")

OK, ERR_BAD_ARG, ERR_BOYS_ORDER, ERR_CUDA, ERR_OOM, ERR_STATE, ERR_UNSUPPORTED = range(7)
_NAMES = {1: "FOCK_ERR_BAD_ARG", 2: "FOCK_ERR_BOYS_ORDER", 3: "FOCK_ERR_CUDA", 4: "FOCK_ERR_OOM", 5: "FOCK_ERR_STATE", 6: "FOCK_ERR_UNSUPPORTED"}

MODE_FAITHFUL, MODE_BOYS_ACCURATE, MODE_FIX_R, MODE_ACCURATE = 0, 1, 2, 3
SCREEN_REFERENCE, SCREEN_OFF, SCREEN_SCHWARZ = 0, 1, 2
STORAGE_INCORE, STORAGE_DIRECT = 0, 1

# every symbol include/fock_b200.h declares
EXPORTS = [
    "fock_default_opts", "fock_create", "fock_eri_build", "fock_eri_get", "fock_eri_export_packed", "fock_packed_len",
    "fock_eri_batch", "fock_build_g", "fock_build_g_device", "fock_one_electron", "fock_set_stream",
    "fock_supermatrix_device", "fock_set_digest_range", "fock_timings", "fock_stats", "fock_last_error", "fock_destroy",
    "fock_abi_version", "fock_measure_fp64_peak", "fock_boys_batch",
    "fock_orbital_eval", "fock_orbital_voxelize", "fock_orbital_stats",
    "fock_pinned_alloc", "fock_pinned_free",
    "fock_diagonal_export", "fock_comm_unique_id", "fock_comm_init", "fock_comm_destroy", "fock_comm_attached", "fock_comm_last_error",
]
COMM_ID_BYTES = 128


class FockError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_NAMES.get(code, code)}: {msg}")
        self.code = code


class CBasis(C.Structure):
    _fields_ = [
        ("n_basis", C.c_int32),
        ("center", C.POINTER(C.c_double)),
        ("angular", C.POINTER(C.c_int32)),
        ("prim_offset", C.POINTER(C.c_int32)),
        ("exponent", C.POINTER(C.c_double)),
        ("coefficient", C.POINTER(C.c_double)),
    ]


class COpts(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("screen", C.c_int32),
        ("tau", C.c_double),
        ("storage", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("keep_eri", C.c_int32),
        ("kernel", C.c_int32),
        ("incore_shard", C.c_int32),
        ("deterministic", C.c_int32),
    ]


_lib = None


def load():
    """Load the shared library or raise -- never falls back to a CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built (no CPU fallback exists). "
            "Run `make -C fock_rs_b200/csrc` or `__graft_entry__.build()`."
        )
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    L.fock_default_opts.argtypes = [C.POINTER(COpts)]
    L.fock_default_opts.restype = None
    L.fock_create.argtypes = [C.POINTER(CBasis), C.POINTER(COpts), C.POINTER(vp)]
    L.fock_eri_build.argtypes = [vp]
    L.fock_eri_get.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp]
    L.fock_eri_export_packed.argtypes = [vp, dp, C.c_size_t]
    L.fock_packed_len.argtypes = [C.c_int32]
    L.fock_packed_len.restype = C.c_size_t
    L.fock_eri_batch.argtypes = [vp, ip, C.c_size_t, dp]
    L.fock_build_g.argtypes = [vp, dp, dp]
    L.fock_build_g_device.argtypes = [vp, vp, vp]
    L.fock_one_electron.argtypes = [vp, C.c_int32, dp, dp, dp, dp, dp]
    L.fock_set_stream.argtypes = [vp, vp]
    L.fock_supermatrix_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.fock_set_digest_range.argtypes = [vp, C.c_size_t, C.c_size_t]
    L.fock_timings.argtypes = [vp, dp, C.c_int32]
    L.fock_stats.argtypes = [vp, C.POINTER(C.c_int64), C.c_int32]
    L.fock_last_error.argtypes = [vp]
    L.fock_last_error.restype = C.c_char_p
    L.fock_destroy.argtypes = [vp]
    L.fock_destroy.restype = None
    L.fock_abi_version.restype = C.c_int
    L.fock_measure_fp64_peak.argtypes = [vp, dp]
    L.fock_boys_batch.argtypes = [vp, dp, C.c_size_t, C.c_int32, dp]
    L.fock_orbital_eval.argtypes = [vp, dp, C.c_double, dp, C.c_size_t, dp]
    L.fock_orbital_voxelize.argtypes = [vp, dp, C.c_double, C.c_int32, C.c_double, C.POINTER(C.c_float)]
    L.fock_orbital_stats.argtypes = [vp, dp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.fock_diagonal_export.argtypes = [vp, dp, C.c_size_t]
    L.fock_comm_unique_id.argtypes = [C.c_char_p]
    L.fock_comm_init.argtypes = [vp, C.c_char_p]
    L.fock_comm_destroy.argtypes = [vp]
    L.fock_comm_attached.argtypes = [vp]
    L.fock_comm_last_error.restype = C.c_char_p
    L.fock_pinned_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.fock_pinned_free.argtypes = [vp]
    if L.fock_abi_version() != 5:
        raise ImportError(f"{LIB_PATH} implements ABI version {L.fock_abi_version()}, this package binds version 5: rebuild it")
    _lib = L
    return L
