#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: tools/summarize_launches.py launches.csv [title] > profiles/xxx.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).strip()
        agg[name][0] += 1
        agg[name][1] += t * scale
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}\n")
    print(f"{n} launches, {tot:.1f} us total device time (ncu per-launch times are cold-cache and serialised: compare SHARES).\n")
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main()
