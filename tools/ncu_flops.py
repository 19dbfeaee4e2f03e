#!/usr/bin/env python3
"""FP64 flops the class kernels of ONE G build actually execute, from an ncu CSV launch list taken with
  --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
(thread-level, predicated-on instruction counts: DFMA = 2 flops, DMUL = DADD = 1).  The launch list may hold several G
builds (warm-up + measured): the totals are divided by the number of times the first class kernel appears.
Prints a markdown table per kernel and, with --json KEY, merges {executed_fp64_flops_per_build, note} into
profiles/ncu_traffic.json under KEY (bench.py reads it for roofline.achieved_executed).

usage: tools/ncu_flops.py launches.csv [--json w32cc/direct/faithful] [--title "..."]
"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = {"dfma": "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "dmul": "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
     "dadd": "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "t": "gpu__time_duration.sum"}


def main():
    path = sys.argv[1]
    key = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else path
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    per = OrderedDict()
    for r in rd:
        name = r.get("Kernel Name", "")
        if "eri_class_kernel" not in name and "eri_class_warpq" not in name:
            continue
        metric, val = r.get("Metric Name"), r.get("Metric Value", "0").replace(",", "")
        k = name.split("(")[0].replace("void ", "").replace("fock::", "")
        d = per.setdefault(k, {"launch_ids": set(), "dfma": 0.0, "dmul": 0.0, "dadd": 0.0, "t": 0.0})
        d["launch_ids"].add(r.get("ID"))
        for short, full in M.items():
            if metric == full:
                v = float(val)
                if short == "t" and r.get("Metric Unit", "").startswith("ns"):
                    v *= 1e-3  # -> us
                elif short == "t" and r.get("Metric Unit", "").startswith("ms"):
                    v *= 1e3
                d[short] += v
    if not per:
        print("no class-kernel rows in", path)
        return
    builds = max(1, min(len(d["launch_ids"]) for d in per.values()))
    tot_flops = tot_t = 0.0
    print(f"# {title}\n")
    print(f"{builds} G build(s) in the capture; figures below are per build.  flops = 2 DFMA + DMUL + DADD (thread-level, predicated on).\n")
    print("| kernel | launches/build | time us (ncu, serialised) | executed GFLOP | DFMA share of FP64 instr | executed TFLOP/s in this launch |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["t"]):
        n = len(d["launch_ids"]) / builds
        fl = (2 * d["dfma"] + d["dmul"] + d["dadd"]) / builds
        t = d["t"] / builds
        inst = d["dfma"] + d["dmul"] + d["dadd"]
        tot_flops += fl
        tot_t += t
        print(f"| `{k}` | {n:.0f} | {t:.1f} | {fl * 1e-9:.2f} | {100 * d['dfma'] / max(inst, 1):.0f} % | {fl / max(t, 1e-9) * 1e-6:.2f} |")
    print(f"\n**Total: {tot_flops:.4e} executed FP64 flops per G build, {tot_t * 1e-3:.2f} ms of serialised class-kernel time "
          f"({tot_flops / max(tot_t, 1e-9) * 1e-6:.2f} TFLOP/s executed).**")
    if key:
        jp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        with open(jp) as f:
            j = json.load(f)
        e = j.setdefault(key, {})
        e["executed_fp64_flops_per_build"] = tot_flops
        e["note"] = f"2*DFMA + DMUL + DADD thread-level predicated-on counts over all class kernels of one G build ({os.path.basename(path)})"
        with open(jp, "w") as f:
            json.dump(j, f, indent=2)
        print(f"\nmerged into profiles/ncu_traffic.json[{key}]")


if __name__ == "__main__":
    main()
