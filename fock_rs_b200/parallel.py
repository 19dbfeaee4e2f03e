"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the exchange.

The path shards naturally: shell-quartet tiles are dealt round-robin to ranks
inside the kernels (`fock_opts.rank/world`), every rank accumulates a partial result, and there is
exactly one exchange step per SCF iteration -- an all-reduce of the N x N partial G.  In-core mode
replicates the one-off build and shards the per-iteration streaming of the packed supermatrix, so each
rank reads only 1/world of it per iteration.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def digest_ranges(n_elements: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous element ranges of the packed supermatrix, one per rank, equal up to padding
    (the layout NCCL reduce-scatter produces)."""
    chunk = (n_elements + world - 1) // world
    return [(min(r * chunk, n_elements), min((r + 1) * chunk, n_elements)) for r in range(world)]


class _DevArray:
    """Minimal __cuda_array_interface__ wrapper so torch can view a raw device buffer."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def device_view(ptr: int, n: int, device=None):
    import torch

    return torch.as_tensor(_DevArray(ptr, n), device=device if device is not None else torch.device("cuda", torch.cuda.current_device()))


def shard_digestion(builder, rank: int, world: int):
    """In-core mode: the (one-off, millisecond-scale) build is replicated on every rank and the
    per-iteration streaming of the packed supermatrix is sharded -- rank r keeps digest_ranges()[r].
    (Sharding the build instead would need a reduce-scatter of the whole packed tensor, 1.9 GB at
    N=208, which costs more than the build.)  fock_eri_build already selects this range by default;
    this helper exists for callers that want a different partition."""
    _, n = builder.supermatrix_device()  # row-padded length
    begin, end = digest_ranges(n, world)[rank]
    builder.set_digest_range(begin, end)


def allreduce_host_matrix(G: np.ndarray) -> np.ndarray:
    """Sum a host matrix over ranks (NCCL needs a device tensor; gloo works on the host)."""
    import torch
    import torch.distributed as dist

    if dist.get_backend() == "nccl":
        t = torch.from_numpy(np.ascontiguousarray(G)).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy().reshape(G.shape)
    t = torch.from_numpy(np.ascontiguousarray(G).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().reshape(G.shape)


def allreduce_device(g_tensor):
    """The per-iteration exchange for device-resident G: one all-reduce of N*N doubles."""
    import torch.distributed as dist

    dist.all_reduce(g_tensor, op=dist.ReduceOp.SUM)
    return g_tensor


class HostStagedG:
    """(Superseded by the collective behind the C ABI -- `FockBuilder.comm_init_torch()` makes `build_g` itself return the
    complete G with one PCIe round trip; kept for hosts that want torch to own the all-reduce.)
    G build for HOST matrices on `world` ranks with ONE PCIe round trip per call: D goes pinned host -> device,
    `fock_build_g_device` produces this rank's partial G, the partials are summed on the device (one NCCL
    all-reduce of N*N doubles) and only the sum returns to the host.  (`fock_build_g` + `allreduce_host_matrix`
    would move G down, up again for NCCL and down once more.)  D must be symmetric, as the reference's density
    is; G is symmetric, so row- and column-major images coincide."""

    def __init__(self, builder, stream=None):
        import torch

        self.b = builder
        n = builder.n
        self.Dp = torch.empty((n, n), dtype=torch.float64).pin_memory()
        self.Gp = torch.empty((n, n), dtype=torch.float64).pin_memory()
        # on the BUILDER's device (fock_opts.device; -1 = the CUDA runtime's current device at fock_create), not torch's current one
        ordinal = int(builder.opts.device)
        dev = torch.device("cuda", ordinal if ordinal >= 0 else torch.cuda.current_device())
        with torch.cuda.device(dev):
            self.stream = stream if stream is not None else torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self.stream):
            self.D = torch.empty((n, n), dtype=torch.float64, device=dev)
            self.G = torch.empty((n, n), dtype=torch.float64, device=dev)
        self.stream.synchronize()

    def __call__(self, D_host) -> np.ndarray:
        import torch
        import torch.distributed as dist

        self.Dp.copy_(torch.from_numpy(np.ascontiguousarray(D_host, dtype=np.float64)))
        with torch.cuda.stream(self.stream):
            self.b.set_stream(self.stream.cuda_stream)
            self.D.copy_(self.Dp, non_blocking=True)
            self.b.build_g_device(self.D.data_ptr(), self.G.data_ptr())
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(self.G, op=dist.ReduceOp.SUM)
            self.Gp.copy_(self.G, non_blocking=True)
        self.stream.synchronize()
        return self.Gp.numpy().copy()
