#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export of one kernel by SOURCE LINE (needs the same -lineinfo build here).

    ncu -i prof.ncu-rep --page source --csv > prof.src.csv
    python tools/ncu_by_line.py prof.src.csv cdftools_b200/libcdfgpu.so mocsig_eos_hist_scan_kernelILb0ELb0 [top]

The ncu export lists SASS instructions in address order; nvdisasm --print-line-info gives the source line of every
instruction of the same kernel in the same order, so the two are joined by position.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_table(so, pat):
    so = os.path.abspath(so)
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    infn, cur, out = False, None, []
    for l in txt.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            infn = pat in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', l)
        if m:
            cur = "%s:%s" % (m.group(1), m.group(2))
            continue
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P\d\s+)?[A-Z0-9_.]+)", l)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2)))
    return out


def main():
    src, so, pat = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = rows[2:]
    lt = line_table(so, pat)
    if len(lt) != len(body):
        print("warning: %d instructions here, %d in the profile -- different build?" % (len(lt), len(body)))
    inst, samp, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for (addr, line, op), r in zip(lt, body):
        n = int(r[ii] or 0)
        inst[line] += n
        samp[line] += int(r[isamp] or 0)
        ops[line][op.split()[-1].split(".")[0]] += n
    tot, tots = sum(inst.values()), sum(samp.values())
    print("instructions executed %d, samples %d" % (tot, tots))
    print("%-28s %12s %6s %8s %6s  top opcodes" % ("line", "inst", "%", "samples", "%"))
    for line, n in inst.most_common(top):
        print("%-28s %12d %6.2f %8d %6.2f  %s" % (line, n, 100.0 * n / tot, samp[line], 100.0 * samp[line] / max(tots, 1),
                                                " ".join("%s:%d" % (k, v * 100 // max(n, 1)) for k, v in ops[line].most_common(4))))


if __name__ == "__main__":
    main()
