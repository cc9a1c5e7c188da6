#!/usr/bin/env python
"""Extract one kernel's SASS from a cubin / .so and print an opcode histogram (build container only; needs cuobjdump).

    python tools/sass_fn.py cdftools_b200/libcdfgpu.so mocsig_eos_hist_scan_kernelILb0ELb0 [--dump out.sass]
"""
import collections
import re
import subprocess
import sys


def main():
    so, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)
    for part in parts[1:]:
        name = part.split("\n", 1)[0].strip()
        if pat not in name:
            continue
        ins = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", part)
        hist = collections.Counter(i.split(".")[0] for i in ins)
        print(name, "instructions:", len(ins))
        print("  " + ", ".join("%s %d" % kv for kv in hist.most_common(40)))
        if "--dump" in sys.argv:
            open(sys.argv[sys.argv.index("--dump") + 1], "w").write(part)


if __name__ == "__main__":
    main()
