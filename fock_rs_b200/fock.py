"""Host-side mirror of the reference's two-electron interface, backed by the CUDA library.

  ElectronTensor.from_basis(basis)  <->  basis/src/electron_tensor.rs:81   (+ `tensor[(i,j,k,l)]`, :179)
  FockBuilder.build_g(D)            <->  scf/src/lib.rs:177-185            (the per-iteration G loop)
  electron_repulsion_int(a,b,c,d)   <->  basis/src/lib.rs:41
  FockEngine                        <->  the three call sites of scf/src/lib.rs:118-185 as one object
                                         (plugs into fock_rs_b200.scf.try_scf)

Everything computes on the GPU through the C ABI (include/fock_b200.h); nothing here has a CPU path.
With torch.distributed initialised (one process per GPU) the quartet tiles are sharded over ranks and
partial G matrices are summed with one all-reduce per iteration.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import FockError
from .basis_set import ContractedGaussian, FlatBasis, flatten


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _as_flat(basis) -> FlatBasis:
    if isinstance(basis, FlatBasis):
        return basis
    return flatten(list(basis))


class FockBuilder:
    """Owns one `fock_ctx`.  Options mirror `fock_opts` (include/fock_b200.h)."""

    def __init__(self, basis, mode=_lib.MODE_FAITHFUL, screen=_lib.SCREEN_REFERENCE, tau=1e-6, storage=_lib.STORAGE_INCORE,
                 device=-1, rank=0, world=1, keep_eri=True, kernel=0, incore_shard=0, deterministic=False):
        self._L = _lib.load()
        self._ctx = C.c_void_p()
        self._pinned = []  # page-locked matrices handed out by pinned_matrix(); they die with the builder
        flat = _as_flat(basis)
        self.flat = flat
        self.n = flat.n_basis
        # keep the arrays alive for the duration of fock_create
        center = np.ascontiguousarray(flat.center, dtype=np.float64)
        angular = np.ascontiguousarray(flat.angular, dtype=np.int32)
        off = np.ascontiguousarray(flat.prim_offset, dtype=np.int32)
        exps = np.ascontiguousarray(flat.exponent, dtype=np.float64)
        coefs = np.ascontiguousarray(flat.coefficient, dtype=np.float64)
        cb = _lib.CBasis(self.n, _dp(center), angular.ctypes.data_as(C.POINTER(C.c_int32)), off.ctypes.data_as(C.POINTER(C.c_int32)), _dp(exps), _dp(coefs))
        opts = _lib.COpts()
        self._L.fock_default_opts(C.byref(opts))
        opts.mode, opts.screen, opts.tau, opts.storage = int(mode), int(screen), float(tau), int(storage)
        opts.device, opts.rank, opts.world, opts.keep_eri, opts.kernel = int(device), int(rank), int(world), int(bool(keep_eri)), int(kernel)
        opts.incore_shard = int(incore_shard)
        opts.deterministic = int(bool(deterministic))
        self.opts = opts
        st = self._L.fock_create(C.byref(cb), C.byref(opts), C.byref(self._ctx))
        if st != _lib.OK:
            msg = self._L.fock_last_error(None)
            self._ctx = C.c_void_p()
            raise FockError(st, msg.decode() if msg else "fock_create failed")

    # -- plumbing
    def _check(self, st):
        if st != _lib.OK:
            msg = self._L.fock_last_error(self._ctx)
            raise FockError(st, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.fock_destroy(self._ctx)
            self._ctx = C.c_void_p()
        for p in getattr(self, "_pinned", []):
            self._L.fock_pinned_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- the path
    def eri_build(self):
        """ElectronTensor::from_basis (+ the electron_terms precompute)."""
        self._check(self._L.fock_eri_build(self._ctx))
        return self

    def eri_get(self, i, j, k, l) -> float:
        v = C.c_double()
        self._check(self._L.fock_eri_get(self._ctx, i, j, k, l, C.byref(v)))
        return v.value

    def eri_export_packed(self) -> np.ndarray:
        out = np.empty(self._L.fock_packed_len(self.n), dtype=np.float64)
        self._check(self._L.fock_eri_export_packed(self._ctx, _dp(out), out.size))
        return out

    def diagonal_export(self) -> np.ndarray:
        """The `diagonal` vector of ElectronTensor::from_basis (basis/src/electron_tensor.rs:90-101): (xy|xy) at tri(x,y)."""
        out = np.empty(self.n * (self.n + 1) // 2, dtype=np.float64)
        self._check(self._L.fock_diagonal_export(self._ctx, _dp(out), out.size))
        return out

    # -- the exchange step behind the ABI (fock_comm_*): after comm_init, build_g returns the complete G on every rank
    @staticmethod
    def comm_unique_id() -> bytes:
        L = _lib.load()
        buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
        st = L.fock_comm_unique_id(buf)
        if st != _lib.OK:
            msg = L.fock_comm_last_error()
            raise FockError(st, msg.decode() if msg else "fock_comm_unique_id failed")
        return buf.raw

    def comm_init(self, unique_id: bytes):
        if len(unique_id) != _lib.COMM_ID_BYTES:
            raise ValueError("the communicator id has FOCK_COMM_ID_BYTES bytes")
        self._check(self._L.fock_comm_init(self._ctx, unique_id))
        return self

    def comm_init_torch(self):
        """Under torch.distributed (any backend): rank 0 creates the id, the process group broadcasts it."""
        import torch.distributed as dist

        box = [self.comm_unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return self.comm_init(box[0])

    @property
    def comm_attached(self) -> bool:
        return bool(self._L.fock_comm_attached(self._ctx))

    def eri_batch(self, quartets) -> np.ndarray:
        q = np.ascontiguousarray(quartets, dtype=np.int32).reshape(-1, 4)
        out = np.empty(q.shape[0], dtype=np.float64)
        self._check(self._L.fock_eri_batch(self._ctx, q.ctypes.data_as(C.POINTER(C.c_int32)), q.shape[0], _dp(out)))
        return out

    def boys(self, t, order: int) -> np.ndarray:
        """`boys::boys(t, n)` (boys/src/lib.rs:58) for an array of arguments, evaluated on the device."""
        tt = np.ascontiguousarray(t, dtype=np.float64).ravel()
        out = np.empty_like(tt)
        self._check(self._L.fock_boys_batch(self._ctx, _dp(tt), tt.size, int(order), _dp(out)))
        return out

    def build_g(self, D, out=None) -> np.ndarray:
        """G = J - K/2 for a symmetric density (host buffers, column-major like DMatrix).  `out`: an N x N column-major float64
        array to receive G; with `D` and `out` from `pinned_matrix()` the call copies straight between them and the device."""
        Dc = np.asfortranarray(D, dtype=np.float64)
        if Dc.shape != (self.n, self.n):
            raise ValueError("density must be N x N")
        if out is None:
            G = np.empty((self.n, self.n), dtype=np.float64, order="F")
        else:
            G = out
            if G.shape != (self.n, self.n) or G.dtype != np.float64 or not G.flags.f_contiguous:
                raise ValueError("out must be an N x N column-major float64 array")
        self._check(self._L.fock_build_g(self._ctx, _dp(Dc), _dp(G)))
        return G

    def pinned_matrix(self) -> np.ndarray:
        """An N x N column-major float64 matrix in page-locked host memory (fock_pinned_alloc), freed with the builder."""
        p = C.c_void_p()
        nbytes = 8 * self.n * self.n
        self._check(self._L.fock_pinned_alloc(nbytes, C.byref(p)))
        self._pinned.append(p)
        buf = (C.c_double * (self.n * self.n)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float64).reshape((self.n, self.n), order="F")

    def build_g_device(self, d_ptr: int, g_ptr: int):
        """Device-resident D and G (raw device pointers, N*N doubles each, column-major)."""
        self._check(self._L.fock_build_g_device(self._ctx, C.c_void_p(d_ptr), C.c_void_p(g_ptr)))

    def one_electron(self, nuc_pos, nuc_charge):
        pos = np.ascontiguousarray(nuc_pos, dtype=np.float64).reshape(-1, 3)
        chg = np.ascontiguousarray(nuc_charge, dtype=np.float64)
        n = self.n
        S, T, V = (np.empty((n, n), dtype=np.float64, order="F") for _ in range(3))
        self._check(self._L.fock_one_electron(self._ctx, pos.shape[0], _dp(pos), _dp(chg), _dp(S), _dp(T), _dp(V)))
        return S, T, V

    def orbital_eval(self, coeff, points, min_coeff: float = 0.05) -> np.ndarray:
        """`MolecularOrbital::evaluate` (scf/src/molecular_orbitals.rs:37-43) for one coefficient column at a
        batch of points; functions with |coeff| <= min_coeff are dropped as `MolecularOrbitals::new` does."""
        c = np.ascontiguousarray(coeff, dtype=np.float64)
        if c.shape != (self.n,):
            raise ValueError("coeff must have N entries")
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.empty(pts.shape[0], dtype=np.float64)
        self._check(self._L.fock_orbital_eval(self._ctx, _dp(c), float(min_coeff), _dp(pts), pts.shape[0], _dp(out)))
        return out

    def orbital_voxelize(self, coeff, n_voxels: int = 200, box_size: float = 20.0, min_coeff: float = 0.05) -> np.ndarray:
        """`State::create_wave_texture_data` (volumetric/src/main.rs:100-125): psi on the n^3 voxel grid as
        f32, x fastest (flat, exactly the R32F texture image the renderer uploads)."""
        c = np.ascontiguousarray(coeff, dtype=np.float64)
        if c.shape != (self.n,):
            raise ValueError("coeff must have N entries")
        out = np.empty(max(int(n_voxels), 0) ** 3, dtype=np.float32)  # a negative size is reported by the C ABI (FOCK_ERR_BAD_ARG)
        self._check(self._L.fock_orbital_voxelize(self._ctx, _dp(c), float(min_coeff), int(n_voxels), float(box_size),
                                                  out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def orbital_stats(self) -> dict:
        ms, nf, ng = C.c_double(), C.c_int32(), C.c_int32()
        self._check(self._L.fock_orbital_stats(self._ctx, C.byref(ms), C.byref(nf), C.byref(ng)))
        return {"kernel_ms": ms.value, "functions": nf.value, "groups": ng.value}

    def set_stream(self, stream_ptr: Optional[int]):
        self._check(self._L.fock_set_stream(self._ctx, C.c_void_p(stream_ptr or 0)))

    def supermatrix_device(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._L.fock_supermatrix_device(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def set_digest_range(self, begin: int, end: int):
        self._check(self._L.fock_set_digest_range(self._ctx, begin, end))

    def measure_fp64_peak(self) -> float:
        """Sustained DFMA rate of this GPU in TFLOP/s (roofline denominator for the ERI kernels)."""
        v = C.c_double()
        self._check(self._L.fock_measure_fp64_peak(self._ctx, C.byref(v)))
        return v.value

    def timings(self) -> dict:
        ms = (C.c_double * 8)()
        self._check(self._L.fock_timings(self._ctx, ms, 8))
        return {"diag_ms": ms[0], "eri_ms": ms[1], "supermatrix_ms": ms[2], "g_device_ms": ms[3], "g_e2e_ms": ms[4],
                "host_sort_ms": ms[5], "host_plan_ms": ms[6], "class_kernels_ms": ms[7]}

    def stats(self) -> dict:
        s = (C.c_int64 * 16)()
        self._check(self._L.fock_stats(self._ctx, s, 16))
        return {
            "n_basis": s[0], "shells": s[1], "shell_pairs": s[2], "shell_quartets_unique": s[3],
            "shell_quartets_evaluated": s[4], "prim_quartets_evaluated": s[5], "launches_build": s[6],
            "launches_g": s[7], "model_flops": s[8], "generic_integrals_evaluated": s[9], "warp_tiles": s[10], "general_shell_quartets_evaluated": s[11],
            "partial_shells": s[12],
        }


class MolecularOrbitals:
    """Host mirror of `scf::molecular_orbitals::MolecularOrbitals` (scf/src/molecular_orbitals.rs:6-100):
    `orbitals[n].evaluate(at)` and the volumetric renderer's voxel grid, evaluated on the device."""

    MIN_COEFFICIENT_MAGNITUDE = 0.05  # molecular_orbitals.rs:47

    class _Orbital:
        def __init__(self, owner, column):
            self._o, self._c = owner, column

        def evaluate(self, at) -> float:
            return float(self._o._b.orbital_eval(self._c, np.asarray(at, dtype=np.float64).reshape(1, 3), self._o.MIN_COEFFICIENT_MAGNITUDE)[0])

        def evaluate_many(self, points) -> np.ndarray:
            return self._o._b.orbital_eval(self._c, points, self._o.MIN_COEFFICIENT_MAGNITUDE)

        def voxelize(self, n_voxels: int = 200, box_size: float = 20.0) -> np.ndarray:
            return self._o._b.orbital_voxelize(self._c, n_voxels, box_size, self._o.MIN_COEFFICIENT_MAGNITUDE)

    def __init__(self, basis, coeff_matrix, builder: Optional[FockBuilder] = None, **opts):
        self._b = builder if builder is not None else FockBuilder(basis, **opts)
        self._C = np.asarray(coeff_matrix, dtype=np.float64)
        if self._C.shape != (self._b.n, self._b.n):
            raise ValueError("coefficient matrix must be N x N")

    def __len__(self):
        return self._C.shape[1]

    def __getitem__(self, n: int) -> "MolecularOrbitals._Orbital":
        return MolecularOrbitals._Orbital(self, np.ascontiguousarray(self._C[:, n]))


class ElectronTensor:
    """`ElectronTensor::from_basis(&basis)` + `tensor[(i,j,k,l)]` (basis/src/electron_tensor.rs:81,179)."""

    def __init__(self, builder: FockBuilder):
        self.builder = builder
        self.size = builder.n

    @classmethod
    def from_basis(cls, basis: Sequence[ContractedGaussian], **opts) -> "ElectronTensor":
        return cls(FockBuilder(basis, **opts).eri_build())

    def __getitem__(self, idx) -> float:
        i, j, k, l = idx
        return self.builder.eri_get(int(i), int(j), int(k), int(l))

    def packed(self) -> np.ndarray:
        return self.builder.eri_export_packed()


def electron_repulsion_int(a: ContractedGaussian, b: ContractedGaussian, c: ContractedGaussian, d: ContractedGaussian, mode=_lib.MODE_FAITHFUL) -> float:
    """`BasisFunction::electron_repulsion_int(a, b, c, d)` (basis/src/lib.rs:41) on the GPU."""
    with FockBuilder([a, b, c, d], mode=mode) as fb:
        return float(fb.eri_batch([[0, 1, 2, 3]])[0])


class FockEngine:
    """The engine `try_scf` drives (see fock_rs_b200.scf).  Rank/world come from torch.distributed
    when it is initialised; partial G matrices are all-reduced over NCCL."""

    def __init__(self, mode=_lib.MODE_FAITHFUL, screen=_lib.SCREEN_REFERENCE, tau=1e-6, storage=_lib.STORAGE_INCORE, device=-1, distributed=None,
                 keep_eri=False, deterministic=False):
        # keep_eri: the SCF loop never reads the raw (ij|kl) tensor (only ElectronTensor / export users do), so the
        # engine does not hold it: half the HBM footprint and build write traffic of an in-core build
        self.kw = dict(mode=mode, screen=screen, tau=tau, storage=storage, device=device, keep_eri=keep_eri, deterministic=deterministic)
        self.builder: Optional[FockBuilder] = None
        self._staged = None
        self.rank, self.world = 0, 1
        if distributed is None:
            try:
                import torch.distributed as dist

                distributed = dist.is_available() and dist.is_initialized()
            except Exception:
                distributed = False
        self.distributed = bool(distributed)
        if self.distributed:
            import torch.distributed as dist

            self.rank, self.world = dist.get_rank(), dist.get_world_size()
            if device < 0:  # one process per GPU: bind to the local rank's device, not whatever happens to be current
                import os

                try:
                    import torch

                    self.kw["device"] = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device())) if torch.cuda.is_available() else -1
                except Exception:
                    pass

    def _ensure(self, flat):
        if self.builder is None or self.builder.flat is not flat:
            if self.builder is not None:
                self.builder.close()
            self.builder = FockBuilder(flat, rank=self.rank, world=self.world, **self.kw)
            if self.distributed and self.world > 1:
                import torch.distributed as dist

                if dist.get_backend() == "nccl":  # the all-reduce then happens inside fock_build_g, on the device
                    self.builder.comm_init_torch()
        return self.builder

    def one_electron(self, flat, nuc_pos, nuc_chg):
        return self._ensure(flat).one_electron(nuc_pos, nuc_chg)

    def from_basis(self, flat):
        b = self._ensure(flat)
        b.eri_build()

    def build_g(self, D):
        if self.distributed and self.world > 1:
            import torch.distributed as dist

            if self.builder.comm_attached:  # one PCIe round trip; partials summed over NVLink inside the C ABI call
                return self.builder.build_g(D)
            from .parallel import allreduce_host_matrix

            return allreduce_host_matrix(self.builder.build_g(D))
        return self.builder.build_g(D)
