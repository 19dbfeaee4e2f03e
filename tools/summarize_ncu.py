#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (read here with `ncu -i`) into a markdown table of the
metrics the roofline discussion uses.  usage: tools/summarize_ncu.py report.ncu-rep [title]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print("`ncu --set full --clock-control none` (each kernel replayed ~40x, cold caches; durations are not bench values).\n")
    cols = [w for w in WANT if w[0] in idx]
    print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("fock::", "")
        vals = []
        for m, _ in cols:
            v, u = r[idx[m]], units[idx[m]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" if abs(f) < 1e4 else f"{f:.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {u}".strip())
        print(f"| `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
