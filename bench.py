#!/usr/bin/env python3
"""bench.py -- ERI + Fock-build benchmark (BASELINE.json metric) for the B200 path and, with
`--impl reference`, for the reference's CPU algorithm on the host cores.

One "step" = one complete pass of the hot path for one SCF iteration's worth of input:
  in-core : ElectronTensor::from_basis (diagonal pass, all surviving shell quartets, supermatrix)
            + one G = J - K/2 build from a density                (nothing cached between steps)
  direct  : one fused ERI -> J/K build from a density (the diagonal pass that provides the screening
            estimates is done once per geometry, outside the step, as in any direct SCF)
`value` = unique shell quartets of the workload / step time (whole job, all ranks), with the density
and the result resident in HBM; `e2e` = the same through the C ABI's host-buffer entry points
(density copied host->device and G device->host inside the timed region every step).

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

WORKLOADS = {
    # name: (description, molecule key, basis)
    "benzene": ("Benzene RHF 6-31G (66 bf)", "benzene", "6-31G"),
    "w16": ("(H2O)16 water cluster 6-31G (208 bf)", "w16", "6-31G"),
    "alkane40": ("Linear alkane C40H82 6-31G (524 bf)", "alkane40", "6-31G"),
    "w32": ("(H2O)32 6-31G** (800 bf Cartesian)", "w32", "6-31G**"),
    "w32cc": ("(H2O)32 cc-pVDZ-shaped synthetic set, general contraction (800 bf Cartesian)", "w32", "cc-pVDZ"),
    "water": ("H2O RHF STO-3G (7 bf)", "water", "STO-3G"),
}


def load_flat(workload):
    from helpers import flat_for

    _, mol, basis = WORKLOADS[workload]
    return flat_for(mol, basis)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
# CPU baseline: the oracle (a C restatement of the reference's algorithm; the Rust reference cannot
# be built in this image) on a bounded sample of the same workload, all host threads.


def count_surviving(diag, thr=1e-6):
    """Number of unordered pairs {xy, zw}, xy != zw, with sqrt(d_xy * d_zw) > thr (the reference's cut),
    counted in O(P log P): with the diagonal sorted descending, the partners of entry i are a prefix."""
    d = np.sort(np.asarray(diag, dtype=np.float64))[::-1]
    d = np.where(np.isfinite(d) & (d > 0), d, 0.0)
    asc = d[::-1]
    total = 0
    with np.errstate(divide="ignore"):
        need = np.where(d > 0, (thr * thr) / d, np.inf)  # d_j must exceed this
    # number of entries strictly greater than need[i] = n - searchsorted(asc, need[i], side='right')
    cnt = d.size - np.searchsorted(asc, need, side="right")
    idx = np.arange(d.size)
    # partners j != i: entry i itself is inside its own prefix iff d_i * d_i > thr^2
    self_in = cnt > idx
    total = int(np.sum(cnt) - np.count_nonzero(self_in))
    return total // 2


def cpu_baseline(workload, screen_ref, budget_s, quartets_unique):
    import oracle

    flat = load_flat(workload)
    ob = oracle.OracleBasis.from_flat(flat)
    n = flat.n_basis
    npair = n * (n + 1) // 2
    flags = oracle.FAITHFUL if screen_ref else oracle.NO_CUT
    threads = oracle.num_threads()
    t0 = time.perf_counter()
    diag = oracle.diagonal(ob, flags, fast=True)  # screening estimates (tabulated evaluator: only to pick the sample)
    n_off_total = npair * (npair - 1) // 2
    n_surv = count_surviving(diag) if screen_ref else n_off_total
    # pilot: a handful of rows to size the sample
    pilot_rows = max(threads, 8)
    stride = max(1, npair // pilot_rows)
    t1 = time.perf_counter()
    nev, _ = oracle.sample_rows(ob, diag, stride // 2, stride, pilot_rows, flags, fast=False)
    dt = time.perf_counter() - t1
    rate = max(nev, 1) / max(dt, 1e-6)
    rows = pilot_rows
    for _ in range(4):  # grow the sample until it costs about budget_s of wall time
        rows = int(min(npair, max(rows + 1, budget_s * rate / max(n_surv / npair, 1.0))))
        stride = max(1, npair // rows)
        t1 = time.perf_counter()
        nev, checksum = oracle.sample_rows(ob, diag, stride // 2, stride, rows, flags, fast=False)
        dt = time.perf_counter() - t1
        rate = max(nev, 1) / max(dt, 1e-6)
        if dt >= 0.6 * budget_s or rows >= npair:
            break
    t_eri = (n_surv + npair) / rate  # + the diagonal pass
    # G build: the reference's N^4/2 loop, timed on a zero tensor of the right size when it fits
    t1 = time.perf_counter()
    if n <= 260:
        packed = np.zeros(oracle.packed_len(n))
        D = np.zeros((n, n))
        oracle.build_g(packed, n, D, parallel=True)
        t_g = time.perf_counter() - t1
        g_note = f"G build timed in full at N={n} (OpenMP, {threads} threads)"
    else:
        m = 120
        packed = np.zeros(oracle.packed_len(m))
        oracle.build_g(packed, m, np.zeros((m, m)), parallel=True)
        t_g = (time.perf_counter() - t1) * (n / m) ** 4
        g_note = f"G build timed at N={m} and scaled by (N/{m})^4"
    t_total = t_eri + t_g
    return {
        "value": quartets_unique / t_total,
        "unit": "shell_quartets/s",
        "cores": threads,
        "kind": "port",
        "sample": (f"{nev} contracted integrals of {n_surv} surviving ({'reference 1e-6 cut' if screen_ref else 'no cut'}) from {rows} evenly "
                   f"strided canonical rows, un-memoised reference recursion, {dt:.1f} s -> {rate:.3e} integrals/s; whole-job ERI time "
                   f"extrapolated to {t_eri:.1f} s; {g_note}: {t_g:.2f} s"),
        "integrals_per_s": rate,
        "eri_s_extrapolated": t_eri,
        "g_s": t_g,
        "s_per_step": t_total,
        "sample_wall_s": time.perf_counter() - t0,
        "checksum": checksum,
    }


def unique_shell_quartets(flat):
    """Unique shell quartets of the workload (s and p of an SP shell are separate shells)."""
    shells = 0
    prev = None
    for i in range(flat.n_basis):
        key = (tuple(flat.center[i]), int(flat.angular[i].sum()), tuple(flat.exponent[flat.prim_offset[i] : flat.prim_offset[i + 1]]))
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
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        budget = 2.0 if i < args.warmup else args.ref_budget
        last = cpu_baseline(args.workload, screen_ref, budget, nq)
        if i >= args.warmup:
            vals.append(last)
    t = float(np.mean([v["s_per_step"] for v in vals]))
    value = nq / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "shell_quartets/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args, flat),
        "cpu_baseline": {k: last[k] for k in ("kind", "cores", "sample")} | {"value": value, "unit": "shell_quartets/s"},
        "e2e": {"value": value, "unit": "shell_quartets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


METRIC = "ERI+Fock-build unique shell quartets/s per SCF iteration (ms_per_step = ERI+Fock-build time per SCF iteration)"


def config_dict(args, flat):
    return {
        "workload": WORKLOADS[args.workload][0],
        "n_basis": flat.n_basis,
        "storage": args.storage,
        "mode": args.mode,
        "screening": "reference 1e-6 cut (basis/src/electron_tensor.rs:134-155)" if args.screen == "ref" else "off",
        "l2": ("inputs larger than L2 (packed supermatrix streamed every step)" if (flat.n_basis >= 150 and args.storage == "incore")
               else "L2 flushed between timed steps (256 MiB write, outside the timed region)"),
        "parallelism": (f"shell-quartet tiles sharded over {args.gpus} rank(s); one all-reduce of G per step"
                        + ("" if args.gpus == 1 or args.storage != "incore" else f"; in-core sharding = {args.incore_shard}")),
    }


def run_ours(args):
    import torch

    from fock_rs_b200 import _lib
    from fock_rs_b200.fock import FockBuilder
    from fock_rs_b200 import parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    flat = load_flat(args.workload)
    n = flat.n_basis
    nq, nshell = unique_shell_quartets(flat)
    mode = _lib.MODE_FAITHFUL if args.mode == "faithful" else _lib.MODE_ACCURATE
    screen = _lib.SCREEN_REFERENCE if args.screen == "ref" else _lib.SCREEN_OFF
    storage = _lib.STORAGE_INCORE if args.storage == "incore" else _lib.STORAGE_DIRECT
    b = FockBuilder(flat, mode=mode, screen=screen, storage=storage, device=local, rank=rank, world=world, keep_eri=False,
                    incore_shard=1 if args.incore_shard == "build" else 0)
    # one explicit (non-default) stream for the library's launches, the NCCL calls and the timing events:
    # torch.cuda.Event only sees the stream it is recorded on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    b.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(20261017)
    Dh = rng.uniform(-1.0, 1.0, (n, n))
    Dh = np.asfortranarray(0.5 * (Dh + Dh.T))
    D = torch.from_numpy(np.ascontiguousarray(Dh.T)).cuda()  # column-major image
    G = torch.empty_like(D)
    needs_flush = n < 150 or storage == _lib.STORAGE_DIRECT  # working set (pair tables, D, y) fits the 126 MB L2
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda") if needs_flush else None

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    direct = storage == _lib.STORAGE_DIRECT
    if direct:
        b.eri_build()  # integral-direct: the diagonal (screening estimates) is per geometry, not per SCF iteration

    def step_device():
        if not direct:
            b.eri_build()
        b.build_g_device(D.data_ptr(), G.data_ptr())
        if world > 1:
            parallel.allreduce_device(G)

    staged = parallel.HostStagedG(b, stream) if world > 1 else None  # N > 1: partials are summed on the device, one PCIe round trip

    def step_host():
        if not direct:
            b.eri_build()
        if staged is not None:
            return staged(Dh)
        return b.build_g(Dh)

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in evs:
            if flush is not None:
                flush.fill_(1.0)
            s.record(stream)
            fn()
            e.record(stream)
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in evs)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    with ClockSampler(local) as clocks:
        barrier()
        t_dev_ms = timed(step_device, args.steps)
        barrier()
        tm = b.timings()
        st = b.stats()
        for _ in range(2):
            step_host()
        barrier()
        t_host_ms = timed(step_host, args.steps)
        barrier()
    g_host = step_host()
    # max over ranks
    if dist is not None:
        t = torch.tensor([t_dev_ms, t_host_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev_ms, t_host_ms = float(t[0]), float(t[1])
    ms_step = t_dev_ms / args.steps
    ms_e2e = t_host_ms / args.steps
    fp64_peak = b.measure_fp64_peak()
    hbm_peak, peak_src = measured_peaks()
    nelem = b._L.fock_packed_len(n)
    # sanity: device and host paths agree
    torch.cuda.synchronize()
    agree = float(np.max(np.abs(G.cpu().numpy().T - g_host)))

    others = []
    if not args.no_other_configs:
        # all ranks take part (sharded runs); rank 0 reports
        for wl, stor in (("benzene", "incore"), ("w16", "incore"), ("alkane40", "direct"), ("w32", "direct"), ("w32cc", "direct")):
            if (wl, stor) != (args.workload, args.storage):
                others.append(quick_config(wl, stor, "ref", rank, world, local, incore_shard=1 if args.incore_shard == "build" else 0))
    if rank == 0:
        # model flops are counted over all ranks; the roofline is per GPU (tiles are dealt evenly), against one GPU's peak
        eri_tf = st["model_flops"] / world / (tm["class_kernels_ms"] * 1e-3) * 1e-12 if tm["class_kernels_ms"] > 0 else 0.0
        streamed = nelem / (world if (world > 1 and args.incore_shard == "digest") else 1)
        g_gbs = 8.0 * streamed / max(tm["g_device_ms"], 1e-9) * 1e-6 if storage == _lib.STORAGE_INCORE else 0.0
        parts = {"diag_ms": tm["diag_ms"] if not direct else 0.0, "eri_ms": tm["eri_ms"] if not direct else 0.0,
                 "supermatrix_ms": tm["supermatrix_ms"], "g_ms": tm["g_device_ms"],
                 "host_sort_ms_inside_eri": tm["host_sort_ms"], "host_plan_ms_inside_eri": tm["host_plan_ms"],
                 "class_kernels_ms_inside_eri": tm["class_kernels_ms"]}
        dominant = max(("diag_ms", "eri_ms", "supermatrix_ms", "g_ms"), key=parts.get)
        if direct:
            dominant = "class_kernels (inside the fused G build)"
        roof_eri = {"bound": "fp64", "achieved": eri_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": eri_tf / fp64_peak if fp64_peak else None,
                    "traffic": ncu_traffic(args.workload, storage == _lib.STORAGE_DIRECT),
                    "kernel": "eri_class_kernel<*> (all class / slice launches of one build, CUDA events around them)",
                    "note": "model flops of the quartets evaluated (SURVEY.md 8d formula; per GPU = total / n_gpus) / time between the events bracketing the class launches; peak = DFMA-chain probe measured in this run (one GPU)"}
        # DRAM bytes of one symv launch from the committed ncu --set full capture (profiles/r01_*_ncu_full_w16_refcut.md)
        symv_traffic = 1.894e9 if (args.workload == "w16" and world == 1 and storage == _lib.STORAGE_INCORE) else None
        roof_g = {"bound": "hbm", "achieved": g_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": g_gbs / hbm_peak, "traffic": symv_traffic,
                  "kernel": "symv_packed_kernel (+ pack/unpack)", "note": f"8 B per stored supermatrix element per rank / device time of one G build; peak = {peak_src}"}
        line = {
            "metric": METRIC, "value": nq / (ms_step * 1e-3), "unit": "shell_quartets/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, flat),
            "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "shell_quartets/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 8 * n * n * world, "d2h_bytes_per_step": 8 * n * n * world},
            "gpu_launches": int((st["launches_g"] if direct else st["launches_build"] + st["launches_g"]) * args.steps),
            "clocks": clocks.summary(),
            "roofline": roof_g if (dominant == "g_ms" and storage == _lib.STORAGE_INCORE) else roof_eri,
            "roofline_eri": roof_eri, "roofline_digest": roof_g if storage == _lib.STORAGE_INCORE else None,
            "breakdown_ms": parts, "dominant": dominant,
            "stats": {"shells": nshell, "shell_quartets_unique": nq, "class_shell_quartets_evaluated": st["shell_quartets_evaluated"],
                      "prim_quartets_evaluated": st["prim_quartets_evaluated"], "model_flops": st["model_flops"], "warp_tiles": st["warp_tiles"], "device_vs_host_path_max_abs_diff": agree},
            "fp64_peak_tflops_measured": fp64_peak,
        }
        line["other_configs"] = others
        if world == 1 and not args.no_other_configs:
            line["orbital_voxelizer"] = voxel_config(local, with_cpu=not args.no_cpu_baseline)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(args.workload, args.screen == "ref", args.cpu_budget, nq).items() if k not in ("checksum",)}
        emit(line)
    b.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def voxel_config(local, with_cpu=True):
    """SURVEY 8 row f4, reported beside the headline: the volumetric renderer's 200^3 x 20 bohr grid
    (volumetric/src/main.rs:63-64,100-125) for one benzene/6-31G orbital through fock_orbital_voxelize
    (host buffer out), and the oracle's restatement of the reference loop on a bounded slab of the same grid."""
    from fock_rs_b200.fock import FockBuilder

    flat = load_flat("benzene")
    rng = np.random.default_rng(20261017)
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

        ob = oracle.OracleBasis.from_flat(flat)
        planes = 4
        lo = 98 * nv * nv
        t0 = time.perf_counter()
        ref = oracle.voxelize(ob, c, nv, box, begin=lo, end=lo + planes * nv * nv)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"kind": "port", "cores": oracle.num_threads(), "sample": f"{planes} of {nv} z-planes", "sample_s": dt,
                               "whole_grid_s_extrapolated": dt * nv / planes, "max_abs_diff_vs_gpu": float(np.max(np.abs(ref - grid[lo:lo + planes * nv * nv])))}
    return out


def ncu_traffic(workload, direct):
    """DRAM bytes (read + write) per launch of the heaviest class kernel of this workload, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written by tools/summarize_ncu.py); None if no
    capture of this workload is on file.  These kernels are FP64-bound: the traffic is megabytes."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload + ("/direct" if direct else "/incore"), {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def quick_config(workload, storage, screen, rank, world, local, steps=5, incore_shard=1):
    """Secondary configs reported beside the headline workload (same step definition, device-resident D/G)."""
    import torch

    from fock_rs_b200 import _lib, parallel
    from fock_rs_b200.fock import FockBuilder

    flat = load_flat(workload)
    n = flat.n_basis
    nq, _ = unique_shell_quartets(flat)
    st_mode = _lib.STORAGE_INCORE if storage == "incore" else _lib.STORAGE_DIRECT
    b = FockBuilder(flat, screen=_lib.SCREEN_REFERENCE if screen == "ref" else _lib.SCREEN_OFF, storage=st_mode, device=local, rank=rank,
                    world=world, keep_eri=False, incore_shard=incore_shard)
    stream = torch.cuda.current_stream()  # run_ours installed an explicit stream as the current one
    assert stream.cuda_stream != 0
    b.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(20261017)
    Dh = rng.uniform(-1.0, 1.0, (n, n))
    D = torch.from_numpy(0.5 * (Dh + Dh.T)).cuda()
    G = torch.empty_like(D)

    if storage == "direct":
        b.eri_build()

    def step():
        if storage != "direct":
            b.eri_build()
        b.build_g_device(D.data_ptr(), G.data_ptr())
        if world > 1:
            parallel.allreduce_device(G)

    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    tm, stt = b.timings(), b.stats()
    b_len = b._L.fock_packed_len(n)
    b.close()
    return {"workload": WORKLOADS[workload][0], "n_basis": n, "storage": storage, "screening": screen, "ms_per_step": ms,
            "shell_quartets_per_s": nq / (ms * 1e-3), "shell_quartets_unique": nq, "shell_quartets_evaluated": stt["shell_quartets_evaluated"],
            "class_kernels_ms": tm["class_kernels_ms"], "g_device_ms": tm["g_device_ms"],
            "digest_hbm_gbs": (8.0 * b_len / (world if incore_shard == 0 else 1) / (tm["g_device_ms"] * 1e-3) * 1e-9) if storage == "incore" else None,
            "class_kernels_model_tflops": stt["model_flops"] / world / (tm["class_kernels_ms"] * 1e-3) * 1e-12 if tm["class_kernels_ms"] > 0 else None}  # per GPU


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
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="alkane40", choices=sorted(WORKLOADS))
    ap.add_argument("--storage", default=None, choices=["incore", "direct"], help="default: direct for alkane40/w32, in-core otherwise")
    ap.add_argument("--screen", default="ref", choices=["ref", "off"])
    ap.add_argument("--mode", default="faithful", choices=["faithful", "accurate"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--ref-budget", type=float, default=12.0, help="seconds of CPU work per step of --impl reference")
    ap.add_argument("--incore-shard", default="build", choices=["build", "digest"],
                    help="N>1, in-core: shard the build (default: the step is 1 build + 1 G build) or replicate it and shard the digestion")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the secondary workloads reported in other_configs")
    args = ap.parse_args()
    if args.storage is None:
        args.storage = "direct" if args.workload in ("alkane40", "w32", "w32cc") else "incore"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
