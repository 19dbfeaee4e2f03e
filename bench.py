#!/usr/bin/env python3
"""bench.py -- ERI + Fock-build benchmark (BASELINE.json metric) for the B200 path and, with
`--impl reference`, for the reference's CPU algorithm on the host cores.

One "step" = one complete pass of the hot path for one SCF iteration's worth of input:
  in-core : ElectronTensor::from_basis (diagonal pass, all surviving shell quartets, supermatrix)
            + one G = J - K/2 build from a density                (nothing cached between steps)
  direct  : one fused ERI -> J/K build from a density (the diagonal pass that provides the screening
            estimates is done once per geometry, outside the step, as in any direct SCF)
`value` = unique shell quartets of the workload / step time (whole job, all ranks), with the density
and the result resident in HBM; `e2e` = the same through the C ABI's host-buffer entry point
`fock_build_g` (density copied host->device and G device->host inside the timed region every step).
With N > 1 ranks the partial G matrices are summed by the NCCL all-reduce INSIDE the C ABI call
(fock_comm_init), in both measurements.

The headline workload is BASELINE.json's largest configuration, config 5: (H2O)32 in the cc-pVDZ-shaped set
(800 Cartesian functions, d shells, 9-primitive contractions), integral-direct, reference cut, the reference's
own arithmetic (faithful mode -- the mode the CPU reference can run).  The other configurations, and config 5
in Boys-accurate mode, are measured the same way and reported with the same keys in `configs`.

Launch:  python bench.py [--gpus N --steps K --warmup W]   (N>1: under torchrun, one rank per GPU)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20261017  # the density of the timed steps = the density of the committed golden G matrices (tests/golden/g_*.npz)

WORKLOADS = {
    # name: (description, molecule key, basis, golden tag stem)
    "benzene": ("Benzene RHF 6-31G (66 bf)", "benzene", "6-31G", None),
    "w16": ("(H2O)16 water cluster 6-31G (208 bf)", "w16", "6-31G", "w16_631g"),
    "alkane40": ("Linear alkane C40H82 6-31G (524 bf)", "alkane40", "6-31G", "alkane40_631g"),
    "w32": ("(H2O)32 6-31G** (800 bf Cartesian)", "w32", "6-31G**", "w32_631gss"),
    "w32cc": ("(H2O)32 cc-pVDZ-shaped synthetic set, general contraction (800 bf Cartesian)", "w32", "cc-pVDZ", "w32_ccpvdz"),
    "water": ("H2O RHF STO-3G (7 bf)", "water", "STO-3G", None),
}
DEFAULT_WORKLOAD = "w32cc"
DIRECT_BY_DEFAULT = ("alkane40", "w32", "w32cc")
# (workload, storage, mode) measured beside the headline with the same keys
# (workload, storage, Boys mode, screening); the last entry is the like-for-like UNSCREENED number SURVEY section 8(d) rule (ii) asks for:
# every unique shell quartet evaluated, no cut, so value = evaluated quartets / s
SECONDARY = [("w32cc", "direct", "accurate", None), ("w32", "direct", "faithful", None), ("alkane40", "direct", "faithful", None),
             ("w16", "incore", "faithful", None), ("benzene", "incore", "faithful", None), ("w16", "direct", "faithful", "off")]
METRIC = "ERI+Fock-build unique shell quartets/s per SCF iteration (ms_per_step = ERI+Fock-build time per SCF iteration)"


def load_flat(workload):
    from helpers import flat_for

    _, mol, basis, _ = WORKLOADS[workload]
    return flat_for(mol, basis)


def seeded_density(n):
    from helpers import random_symmetric_density

    return random_symmetric_density(n, seed=SEED)


def golden_g(workload, mode):
    """The oracle's G for the seeded density at full size (tools/make_golden_large.py), or None."""
    stem = WORKLOADS[workload][3]
    if stem is None:
        return None
    path = os.path.join(ROOT, "tests", "golden", f"g_{stem}_{mode}.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    if int(z["seed"][0]) != SEED:
        return None
    n = load_flat(workload).n_basis
    G = np.zeros((n, n))
    G[np.triu_indices(n)] = z["g_upper"]
    return G + np.triu(G, 1).T


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML (spawning
    nvidia-smi every 100 ms perturbs a 3 ms step measurably); nvidia-smi is the fallback."""

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.samples = []  # (sm_mhz, reason bitmask)
        self._stop = threading.Event()
        self._t = None
        self._nv = None
        self._h = None
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM)  # first call is slow
        except Exception:
            self._h = None

    def _sample_nvml(self):
        nv = self._nv
        mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        self.samples.append((float(mhz), int(reasons)))

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(",")]
        bits = 0
        for i, b in enumerate((0x8, 0x40, 0x20, 0x4)):
            if len(f) > 2 + i and f[2 + i].lower().startswith("active"):
                bits |= b
        self.max_mhz = float(f[1])
        self.samples.append((float(f[0]), bits))

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.003 if self._h is not None else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        bits = 0
        for _, r in self.samples:
            bits |= r
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": [n for b, n in self.NAMES.items() if bits & b],
                "samples": len(self.samples), "source": "nvml" if self._h is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C restatement of the reference's algorithm; the Rust reference cannot be built in
# this image) on all host threads.  The job is the reference's tensor build + G loop in integral-direct form
# (oracle_direct_g: basis/src/electron_tensor.rs:113-158 visited slot by slot with the un-memoised recursions
# of basis/src/utils.rs, each integral digested by the definition of scf/src/lib.rs:167-185) -- the only form
# that can run at all beyond N ~ 250, where the reference's dense N^4 tensor no longer fits any host.


def count_surviving(diag, thr=1e-6):
    """Number of unordered pairs {xy, zw}, xy != zw, with sqrt(d_xy * d_zw) > thr (the reference's cut),
    counted in O(P log P): with the diagonal sorted descending, the partners of entry i are a prefix."""
    d = np.sort(np.where(np.isfinite(diag) & (diag > 0), diag, 0.0))[::-1]
    n = d.size
    asc = d[::-1]
    need = np.full(n, np.inf)
    np.divide(thr * thr, d, out=need, where=d > thr * thr * 1e-300)
    cnt = n - np.searchsorted(asc, need, side="right")
    for _ in range(4):  # settle each prefix boundary with the reference's exact expression
        keep = (cnt == 0) | (np.sqrt(d * d[np.maximum(cnt - 1, 0)]) > thr)
        cnt = np.where(keep, cnt, cnt - 1)
        more = (cnt < n) & (np.sqrt(d * d[np.minimum(cnt, n - 1)]) > thr)
        cnt = np.where(more, cnt + 1, cnt)
    self_in = cnt > np.arange(n)
    return int(np.sum(cnt) - np.count_nonzero(self_in)) // 2


_CPU_SETUP = {}


def cpu_setup(workload, screen_ref):
    """Per-workload inputs of the CPU arm, computed once per process: basis, density, the screening diagonal
    (tabulated evaluator -- it only decides WHICH integrals the timed loop evaluates) and the survivor count."""
    key = (workload, screen_ref)
    if key not in _CPU_SETUP:
        import oracle

        oracle.set_num_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core it may
        flat = load_flat(workload)
        ob = oracle.OracleBasis.from_flat(flat)
        n = flat.n_basis
        npair = n * (n + 1) // 2
        flags = oracle.FAITHFUL if screen_ref else oracle.NO_CUT
        diag = oracle.diagonal(ob, flags, fast=True)
        n_surv = count_surviving(diag) if screen_ref else npair * (npair - 1) // 2
        _CPU_SETUP[key] = (flat, ob, n, npair, flags, diag, n_surv, seeded_density(n))
    return _CPU_SETUP[key]


def cpu_sample(workload, screen_ref, budget_s, quartets_unique, rate_hint=None):
    """One bounded sample of the CPU job: evenly strided canonical rows through oracle_direct_g (un-memoised
    recursion = what the reference executes), sized for ~budget_s of wall time; the whole job when it fits."""
    import oracle

    flat, ob, n, npair, flags, diag, n_surv, D = cpu_setup(workload, screen_ref)
    threads = oracle.num_threads()
    total = n_surv + npair
    t0 = time.perf_counter()
    rate = rate_hint
    if rate is None:  # pilot: one row per thread, evenly strided
        rows = max(threads, 8)
        stride = max(1, npair // rows)
        t1 = time.perf_counter()
        _, nev = oracle.direct_g(ob, D, flags, fast=False, diag=diag, row_start=stride // 2, row_stride=stride, max_rows=rows)
        rate = max(nev, 1) / max(time.perf_counter() - t1, 1e-6)
    rows = int(min(npair, max(threads, budget_s * rate / max(total / npair, 1.0))))
    stride = max(1, npair // rows)
    full = stride == 1
    t1 = time.perf_counter()
    _, nev = oracle.direct_g(ob, D, flags, fast=False, diag=diag, row_start=0 if full else stride // 2, row_stride=stride,
                             max_rows=npair if full else rows)
    dt = time.perf_counter() - t1
    rate = max(nev, 1) / max(dt, 1e-6)
    frac = min(1.0, nev / total)
    t_step = dt if full else total / rate
    return {
        "value": quartets_unique / t_step, "unit": "shell_quartets/s", "cores": threads, "kind": "port",
        "sample": (f"{f'all {nev}' if full else f'{nev} of {total}'} integrals ({'reference 1e-6 cut' if screen_ref else 'no cut'}; "
                   f"{npair if full else rows} {'' if full else 'evenly strided '}canonical rows of {npair}) evaluated with the reference's un-memoised "
                   f"recursion and digested into G (integral-direct restatement of electron_tensor.rs:113-158 + scf/lib.rs:167-185), "
                   f"{dt:.1f} s on {threads} threads = {rate:.3e} integrals/s"
                   + ("" if full else f"; whole-job time extrapolated by 1/{frac:.4f} to {t_step:.1f} s")),
        "extrapolated": not full, "sample_fraction": frac, "integrals_per_s": rate, "s_per_step": t_step, "sample_wall_s": time.perf_counter() - t0,
    }


def cpu_g_loop(workload):
    """The reference's own per-iteration G loop (scf/src/lib.rs:177-185: dense N^4/2 FMAs over `electron_terms`) at the REAL N,
    serial as in the reference and with OpenMP -- only where the packed tensor fits the host (N <= 208: 1.9 GB)."""
    import oracle

    flat = load_flat(workload)
    n = flat.n_basis
    if n > 260:
        return {"n_basis": n, "note": "the reference's in-core G loop needs its dense N^4 tensor (603 GB at N=524, 3.3 TB at N=800): it cannot run at this size; "
                                      "the CPU arm is the integral-direct restatement, whose digestion is inside the timed loop"}
    packed = np.zeros(oracle.packed_len(n))
    D = seeded_density(n)
    t0 = time.perf_counter()
    oracle.build_g(packed, n, D, parallel=True)
    t_par = time.perf_counter() - t0
    out = {"n_basis": n, "openmp_s": t_par, "threads": oracle.num_threads()}
    t0 = time.perf_counter()
    oracle.build_g(packed, n, D, parallel=False)  # scf/src/utils.rs:25-28 is serial: N^4/2 FMAs on one core
    out["serial_s"] = time.perf_counter() - t0
    return out


def unique_shell_quartets(flat):
    """Unique shell quartets of the workload (s and p of an SP shell are separate shells)."""
    shells = 0
    prev = None
    for i in range(flat.n_basis):
        key = (tuple(flat.center[i]), int(flat.angular[i].sum()), tuple(flat.exponent[flat.prim_offset[i] : flat.prim_offset[i + 1]]),
               tuple(flat.coefficient[flat.prim_offset[i] : flat.prim_offset[i + 1]]) if int(flat.angular[i].sum()) == 0 else None)
        if key != prev:
            shells += 1
            prev = key
    sp = shells * (shells + 1) // 2
    return sp * (sp + 1) // 2, shells


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    flat = load_flat(args.workload)
    nq, _ = unique_shell_quartets(flat)
    screen_ref = args.screen == "ref"
    # the driver runs this arm with its own --steps/--warmup: size each step's sample so the whole run stays within ~4 minutes
    n_steps = args.warmup + args.steps
    budget = args.ref_budget if args.ref_budget > 0 else max(2.0, min(20.0, 200.0 / n_steps))
    t_all = time.perf_counter()
    vals, last, rate = [], None, None
    for i in range(n_steps):
        last = cpu_sample(args.workload, screen_ref, budget if i >= args.warmup else min(budget, 3.0), nq, rate_hint=rate)
        rate = last["integrals_per_s"]
        if i >= args.warmup:
            vals.append(last)
    t = float(np.mean([v["s_per_step"] for v in vals]))
    value = nq / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "shell_quartets/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args, flat),
        "extrapolated": bool(last["extrapolated"]), "sample_fraction": last["sample_fraction"],
        "measured_wall_s": time.perf_counter() - t_all,
        "cpu_baseline": {k: last[k] for k in ("kind", "cores", "sample", "extrapolated", "sample_fraction", "integrals_per_s")} | {"value": value, "unit": "shell_quartets/s"},
        "reference_g_loop": cpu_g_loop(args.workload),
        "e2e": {"value": value, "unit": "shell_quartets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def config_dict(args, flat, workload=None, storage=None, mode=None, screen=None):
    workload = workload or args.workload
    storage = storage or args.storage
    mode = mode or args.mode
    screen = screen or args.screen
    return {
        "workload": WORKLOADS[workload][0],
        "n_basis": flat.n_basis,
        "storage": storage,
        "mode": mode,
        "screening": "reference 1e-6 cut (basis/src/electron_tensor.rs:134-155)" if screen == "ref" else "off (every unique shell quartet evaluated)",
        "l2": ("inputs larger than L2 (packed supermatrix streamed every step)" if (flat.n_basis >= 150 and storage == "incore")
               else "L2 flushed between timed steps (256 MiB write, outside the timed region)"),
        "parallelism": (f"shell-quartet tiles sharded over {args.gpus} rank(s); one NCCL all-reduce of G per step inside fock_build_g"
                        + ("" if args.gpus == 1 or storage != "incore" else f"; in-core sharding = {args.incore_shard}")),
    }


class Env:
    """Process-wide GPU state of one bench run: rank/world, the one explicit stream, the L2 flush buffer."""

    def __init__(self):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        # one explicit (non-default) stream for the library's launches, its NCCL call and the timing events:
        # torch.cuda.Event only sees the stream it is recorded on
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0
        self.flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return vals
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return tuple(float(x) for x in t)


def bench_config(env, args, workload, storage, mode, steps, warmup, clocks_for=None, screen=None):
    """Measures one configuration: device-resident steps, host-buffer (e2e) steps, roofline, parity against the golden G."""
    screen = screen or args.screen
    torch = env.torch
    from fock_rs_b200 import _lib
    from fock_rs_b200.fock import FockBuilder

    flat = load_flat(workload)
    n = flat.n_basis
    nq, nshell = unique_shell_quartets(flat)
    c_mode = _lib.MODE_FAITHFUL if mode == "faithful" else _lib.MODE_ACCURATE
    c_screen = _lib.SCREEN_REFERENCE if screen == "ref" else _lib.SCREEN_OFF
    direct = storage == "direct"
    b = FockBuilder(flat, mode=c_mode, screen=c_screen, storage=_lib.STORAGE_DIRECT if direct else _lib.STORAGE_INCORE, device=env.local,
                    rank=env.rank, world=env.world, keep_eri=False, incore_shard=1 if args.incore_shard == "build" else 0)
    b.set_stream(env.stream.cuda_stream)
    if env.world > 1:
        b.comm_init_torch()  # from here on fock_build_g[_device] returns the complete G on every rank
    Dh = np.asfortranarray(seeded_density(n))
    D = torch.from_numpy(np.ascontiguousarray(Dh.T)).cuda()  # column-major image
    G = torch.empty_like(D)
    needs_flush = n < 150 or direct  # working set (pair tables, D, A) fits the 126 MB L2
    if direct:
        b.eri_build()  # integral-direct: the diagonal (screening estimates) is per geometry, not per SCF iteration

    def step_device():
        if not direct:
            b.eri_build()
        b.build_g_device(D.data_ptr(), G.data_ptr())

    # host-buffer path: the caller's D and G are page-locked matrices (fock_pinned_alloc), as the bench contract's e2e asks --
    # fock_build_g copies them straight to / from the device
    Dpin, Gpin = b.pinned_matrix(), b.pinned_matrix()
    Dpin[...] = Dh

    def step_host():
        if not direct:
            b.eri_build()
        return b.build_g(Dpin, out=Gpin)

    def timed(fn, k):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        for s, e in evs:
            if needs_flush:
                env.flush.fill_(1.0)
            s.record(env.stream)
            fn()
            e.record(env.stream)
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in evs)

    for _ in range(max(warmup, 3)):
        step_device()
    env.barrier()
    sampler = ClockSampler(env.local) if clocks_for is not None else None
    if sampler:
        sampler.__enter__()
    env.barrier()
    t_dev_ms = timed(step_device, steps)
    env.barrier()
    tm, st = b.timings(), b.stats()
    for _ in range(2):
        step_host()
    env.barrier()
    t_host_ms = timed(step_host, steps)
    env.barrier()
    if sampler:
        sampler.__exit__()
        clocks_for.update(sampler.summary())
    g_host = step_host().copy()
    t_dev_ms, t_host_ms = env.max_over_ranks(t_dev_ms, t_host_ms)
    ms_step, ms_e2e = t_dev_ms / steps, t_host_ms / steps
    torch.cuda.synchronize()
    agree = float(np.max(np.abs(G.cpu().numpy().T - g_host)))
    want = golden_g(workload, mode) if screen == "ref" else None
    parity = None if want is None else float(np.max(np.abs(g_host - want)))
    fp64_peak = b.measure_fp64_peak()
    hbm_peak, peak_src = measured_peaks()
    # In-core configurations, the shape of a real SCF run: ONE tensor build, then a G build per iteration that only streams the
    # packed supermatrix.  At N > 1 the build is replicated and the streaming is sharded by contiguous element ranges
    # (incore_shard = 0), the partial G summed by the all-reduce behind fock_build_g_device -- the regime that scales.
    scf_shaped = None
    if not direct:
        k_iter = 15
        b2 = FockBuilder(flat, mode=c_mode, screen=c_screen, storage=_lib.STORAGE_INCORE, device=env.local, rank=env.rank, world=env.world,
                         keep_eri=False, incore_shard=0)
        b2.set_stream(env.stream.cuda_stream)
        if env.world > 1:
            b2.comm_init_torch()
        b2.eri_build()
        b2.build_g_device(D.data_ptr(), G.data_ptr())
        env.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(env.stream)
        b2.eri_build()
        s1.record(env.stream)
        torch.cuda.synchronize()
        build_ms = s0.elapsed_time(s1)
        env.barrier()
        g_total = timed(lambda: b2.build_g_device(D.data_ptr(), G.data_ptr()), k_iter)
        env.barrier()
        build_ms, g_total = env.max_over_ranks(build_ms, g_total)
        g2 = G.cpu().numpy().T.copy()
        scf_shaped = {"tensor_builds": 1, "g_builds": k_iter, "build_ms": build_ms, "ms_per_g_build": g_total / k_iter,
                      "total_ms": build_ms + g_total, "sharding": "build replicated, supermatrix streaming sharded over ranks" if env.world > 1 else "one GPU",
                      "g_vs_first_path_max_abs_diff": float(np.max(np.abs(g2 - g_host)))}
        b2.close()
    nelem = b._L.fock_packed_len(n)
    out = None
    if env.rank == 0:
        world = env.world
        eri_tf = st["model_flops"] / world / (tm["class_kernels_ms"] * 1e-3) * 1e-12 if tm["class_kernels_ms"] > 0 else 0.0
        streamed = nelem / (world if (world > 1 and args.incore_shard == "digest") else 1)
        g_gbs = 8.0 * streamed / max(tm["g_device_ms"], 1e-9) * 1e-6 if not direct else 0.0
        parts = {"diag_ms": tm["diag_ms"] if not direct else 0.0, "eri_ms": tm["eri_ms"] if not direct else 0.0,
                 "supermatrix_ms": tm["supermatrix_ms"], "g_ms": tm["g_device_ms"],
                 "host_sort_ms_inside_eri": tm["host_sort_ms"], "host_plan_ms_inside_eri": tm["host_plan_ms"],
                 "class_kernels_ms_inside_eri_or_g": tm["class_kernels_ms"]}
        dominant = "class_kernels (inside the fused G build)" if direct else max(("diag_ms", "eri_ms", "supermatrix_ms", "g_ms"), key=parts.get)
        ncu = ncu_record(workload, direct, mode)
        roof_eri = {"bound": "fp64", "achieved": eri_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": eri_tf / fp64_peak if fp64_peak else None,
                    "traffic": ncu.get("dram_bytes_per_launch"),
                    "kernel": "eri_class_kernel<*> (all class / slice launches of one build, CUDA events around them)",
                    "note": "achieved = MODEL flops of the quartets evaluated (SURVEY.md 8d formula; per GPU = total / n_gpus) / time between the events bracketing "
                            "the class launches; peak = DFMA-chain probe measured in this run (one GPU; MEASURED_PEAKS.json has no FP64 figure)"}
        if ncu.get("executed_fp64_flops_per_build"):
            # the model credits every primitive quartet with the near-field Boys cost; far-field ones execute far fewer flops
            ex = ncu["executed_fp64_flops_per_build"] / world / (tm["class_kernels_ms"] * 1e-3) * 1e-12
            roof_eri["executed_fp64_flops_per_build"] = ncu["executed_fp64_flops_per_build"]
            roof_eri["achieved_executed"] = ex
            roof_eri["frac_executed"] = ex / fp64_peak if fp64_peak else None
            roof_eri["executed_note"] = ncu.get("note")
        roof_g = {"bound": "hbm", "achieved": g_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": g_gbs / hbm_peak,
                  "traffic": 1.894e9 if (workload == "w16" and world == 1 and not direct) else None,
                  "kernel": "symv_packed_kernel (+ pack/unpack)", "note": f"8 B per stored supermatrix element per rank / device time of one G build; peak = {peak_src}"}
        out = {
            "metric": METRIC, "value": nq / (ms_step * 1e-3), "unit": "shell_quartets/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, flat, workload, storage, mode, screen),
            "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "shell_quartets/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 8 * n * n * world, "d2h_bytes_per_step": 8 * n * n * world},
            "gpu_launches": int((st["launches_g"] if direct else st["launches_build"] + st["launches_g"]) * steps),
            "roofline": roof_g if (dominant == "g_ms" and not direct) else roof_eri,
            "roofline_eri": roof_eri, "roofline_digest": roof_g if not direct else None,
            "breakdown_ms": parts, "dominant": dominant,
            "parity": {"g_vs_oracle_max_abs_diff": parity, "tolerance": 1e-9,
                       "oracle": "tests/golden/g_*.npz (CPU oracle in integral-direct form, same seeded density)" if parity is not None else "no full-size fixture for this configuration",
                       "device_vs_host_path_max_abs_diff": agree},
            "stats": {"shells": nshell, "shell_quartets_unique": nq, "class_shell_quartets_evaluated": st["shell_quartets_evaluated"],
                      "evaluated_shell_quartets_per_s": st["shell_quartets_evaluated"] / (ms_step * 1e-3),
                      "prim_quartets_evaluated": st["prim_quartets_evaluated"], "model_flops": st["model_flops"], "warp_tiles": st["warp_tiles"]},
            "fp64_peak_tflops_measured": fp64_peak,
        }
        if scf_shaped is not None:
            out["scf_shaped"] = scf_shaped
    b.close()
    return out


def run_ours(args):
    env = Env()
    clocks = {}
    line = bench_config(env, args, args.workload, args.storage, args.mode, args.steps, args.warmup, clocks_for=clocks)
    others = []
    if not args.no_other_configs:
        for wl, stor, mode, scr in SECONDARY:  # all ranks take part (sharded runs); rank 0 reports
            if (wl, stor, mode, scr or args.screen) != (args.workload, args.storage, args.mode, args.screen):
                others.append(bench_config(env, args, wl, stor, mode, args.secondary_steps, 3, screen=scr))
    if env.rank == 0:
        line["clocks"] = clocks
        line["configs"] = others
        if env.world == 1 and not args.no_other_configs:
            line["orbital_voxelizer"] = voxel_config(env.local, with_cpu=not args.no_cpu_baseline)
        if env.world == 1 and not args.no_cpu_baseline:
            nq = line["stats"]["shell_quartets_unique"]
            line["cpu_baseline"] = cpu_sample(args.workload, args.screen == "ref", args.cpu_budget, nq)
        emit(line)
    if env.dist is not None:
        env.dist.barrier()
        env.dist.destroy_process_group()


def voxel_config(local, with_cpu=True):
    """SURVEY 8 row f4, reported beside the headline: the volumetric renderer's 200^3 x 20 bohr grid
    (volumetric/src/main.rs:63-64,100-125) for one benzene/6-31G orbital through fock_orbital_voxelize
    (host buffer out), and the oracle's restatement of the reference loop on a bounded slab of the same grid."""
    from fock_rs_b200.fock import FockBuilder

    flat = load_flat("benzene")
    rng = np.random.default_rng(SEED)
    c = rng.uniform(-1.0, 1.0, flat.n_basis)
    nv, box = 200, 20.0
    with FockBuilder(flat, device=local) as b:
        b.orbital_voxelize(c, nv, box)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            grid = b.orbital_voxelize(c, nv, box)
        e2e_ms = (time.perf_counter() - t0) / reps * 1e3
        st = b.orbital_stats()
    nprim = int(sum(flat.prim_offset[i + 1] - flat.prim_offset[i] for i in range(flat.n_basis) if abs(c[i]) > 0.05))
    out = {"workload": "benzene/6-31G orbital on the renderer's 200^3 voxel grid (box 20 bohr)", "voxels": nv ** 3, "functions": st["functions"],
           "exp_groups": st["groups"], "kernel_ms": st["kernel_ms"], "e2e_ms_host_grid_out": e2e_ms, "d2h_bytes": 4 * nv ** 3,
           "voxels_per_s": nv ** 3 / (st["kernel_ms"] * 1e-3), "reference_primitive_terms_per_s": nprim * nv ** 3 / (st["kernel_ms"] * 1e-3)}
    if with_cpu:
        import oracle

        oracle.set_num_threads(host_threads())
        ob = oracle.OracleBasis.from_flat(flat)
        planes = 4
        lo = 98 * nv * nv
        t0 = time.perf_counter()
        ref = oracle.voxelize(ob, c, nv, box, begin=lo, end=lo + planes * nv * nv)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"kind": "port", "cores": oracle.num_threads(), "sample": f"{planes} of {nv} z-planes", "sample_s": dt,
                               "whole_grid_s_extrapolated": dt * nv / planes, "max_abs_diff_vs_gpu": float(np.max(np.abs(ref - grid[lo:lo + planes * nv * nv])))}
    return out


def ncu_record(workload, direct, mode):
    """What the committed ncu captures hold for this configuration (profiles/ncu_traffic.json, written by
    tools/summarize_ncu.py / tools/ncu_flops.py): DRAM bytes per launch of the heaviest class kernel (`--set full`) and
    the FP64 flops the class kernels of one build actually execute (smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on).
    {} if no capture of this configuration is on file.  These kernels are FP64-bound: the traffic is megabytes."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        key = workload + ("/direct" if direct else "/incore")
        rec = dict(d.get(key, {}) or {})
        rec.update(d.get(key + "/" + mode, {}) or {})
        return rec
    except (OSError, ValueError):
        return {}


_JSON_FD = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything libraries print while the bench runs
    (NCCL's version banner, for one) has been routed to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--storage", default=None, choices=["incore", "direct"], help="default: direct for alkane40/w32/w32cc, in-core otherwise")
    ap.add_argument("--screen", default="ref", choices=["ref", "off"])
    ap.add_argument("--mode", default="faithful", choices=["faithful", "accurate"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--ref-budget", type=float, default=0.0, help="seconds of CPU work per step of --impl reference (0: sized so the whole run takes ~4 minutes)")
    ap.add_argument("--secondary-steps", type=int, default=5, help="timed steps of each configuration in `configs`")
    ap.add_argument("--incore-shard", default="build", choices=["build", "digest"],
                    help="N>1, in-core: shard the build (default: the step is 1 build + 1 G build) or replicate it and shard the digestion")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the secondary workloads reported in `configs`")
    args = ap.parse_args()
    if args.storage is None:
        args.storage = "direct" if args.workload in DIRECT_BY_DEFAULT else "incore"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
