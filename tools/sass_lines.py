#!/usr/bin/env python
"""Per-source-line SASS statistics of one kernel (build container only; needs cuobjdump + nvdisasm, -lineinfo build).

    python tools/sass_lines.py cdftools_b200/libcdfgpu.so mocsig_eos_hist_scan_kernelILb0ELb0 [--ops STL,LDL]
Prints, per source line, the number of SASS instructions and how many of them match --ops.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    so, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
    ops = tuple(sys.argv[sys.argv.index("--ops") + 1].split(",")) if "--ops" in sys.argv else ("STL", "LDL")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    infn, cur = False, None
    cnt, tot = collections.Counter(), collections.Counter()
    for l in txt.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            infn = pat in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m and cur:
            tot[cur] += 1
            if m.group(1).startswith(ops):
                cnt[cur] += 1
    print("total instructions %d, matching %d" % (sum(tot.values()), sum(cnt.values())))
    for k, v in sorted(tot.items()):
        if "--all" in sys.argv or cnt[k]:
            print("%-26s %5d  %5d" % ("%s:%d" % k, v, cnt[k]))


if __name__ == "__main__":
    main()
