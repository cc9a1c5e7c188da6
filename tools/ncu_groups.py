#!/usr/bin/env python
"""Instruction / sample shares of K2's functions from an ncu source-page export (see tools/ncu_by_line.py).

    python tools/ncu_groups.py prof.src.csv prof.raw.csv [windows]
"""
import collections
import csv
import re
import sys

sys.path.insert(0, "tools")
import ncu_by_line as n

SO, KERNEL, SRC = "cdftools_b200/libcdfgpu.so", "mocsig_eos_hist_scan_kernelILb0ELb0", "cdftools_b200/csrc/mocsig_kernel.cuh"


def function_ranges():
    """(first line, name) of every function / struct in the kernel source, in order"""
    out = []
    lines = open(SRC).read().split("\n")
    for i, l in enumerate(lines, 1):
        m = re.match(r"(?:__device__|__global__).*?\b(\w+)\(", l)
        if m and not l.startswith(" "):
            out.append((i, m.group(1)))
    return out


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, body = rows[1], rows[2:]
    ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    lt = n.line_table(SO, KERNEL)
    fr = function_ranges()
    g, gs = collections.Counter(), collections.Counter()
    for (addr, line, op), r in zip(lt, body):
        f, l = (line or ":0").split(":")
        name = f
        if f == "mocsig_kernel.cuh":
            name = "?"
            for first, fn in fr:
                if first <= int(l) + 2:
                    name = fn
        g[name] += int(r[ii] or 0)
        gs[name] += int(r[isamp] or 0)
    tot, ts = sum(g.values()), sum(gs.values())
    win = float(sys.argv[3]) if len(sys.argv) > 3 else 453324.0
    print("%-34s %12s %6s %8s %10s" % ("function", "inst", "%", "samples%", "per window"))
    for k, v in g.most_common(16):
        print("%-34s %12d %5.1f%% %7.1f%% %10.1f" % (k, v, 100 * v / tot, 100 * gs[k] / max(ts, 1), v / win))
    print("%-34s %12d" % ("total", tot))
    if len(sys.argv) > 2:
        raw = list(csv.reader(open(sys.argv[2])))
        for h, v, u in zip(raw[0], raw[2], raw[1]):
            if h in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                     "dram__bytes_read.sum", "dram__bytes_write.sum") or (h.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in h and float(v or 0) > 0.2):
                print("  %s = %s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall "), v, u))


if __name__ == "__main__":
    main()
