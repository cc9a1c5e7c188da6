#!/bin/bash
# by-source-line instruction / stall table of one kernel of libcdfgpu.so from an ncu report:
#   tools/ncu_lines.sh report.ncu-rep <mangled-name-substring> [ntop] [lib.so]
set -e
rep=$1; sym=$2; ntop=${3:-50}; lib=${4:-cdftools_b200/libcdfgpu.so}
tmp=$(mktemp -d)
( cd $tmp && cuobjdump -xelf all $OLDPWD/$lib >/dev/null 2>&1 && nvdisasm -g *.cubin > all.txt 2>/dev/null )
python - "$tmp/all.txt" "$sym" > $tmp/lines.txt <<'PY'
import sys,re
on=False
for ln in open(sys.argv[1]):
    m=re.match(r'\s*\.section\s+\.text\.(\S+?),',ln)
    if m: on = sys.argv[2] in m.group(1)
    if on: sys.stdout.write(ln)
PY
ncu -i $rep --page source --csv 2>/dev/null > $tmp/src.csv
python tools/ncu_by_line.py $tmp/src.csv $tmp/lines.txt $ntop
rm -rf $tmp
