"""Join an `ncu --page source --csv` dump (per-SASS-instruction counters) with `nvdisasm -g` line info of the same
cubin, and aggregate executed warp-instructions and stall samples per CUDA source line.
usage: python tools/ncu_by_line.py ncu_source.csv nvdisasm_lines.txt [ntop]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ie, ns = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ti = hdr.index('Thread Instructions Executed')
insts = []
for r in rows[2:]:
    try: insts.append((int(r[ia], 16), int(r[ie]), int(r[ns]), int(r[ti])))
    except Exception: pass
base = insts[0][0]
line_of = {}
cur = None
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m: line_of[int(m.group(1), 16)] = cur
import os
SORTKEY = 1 if os.environ.get("BY_SAMPLES") else 0
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for a, n, s, t in insts:
    k = line_of.get(a - base)
    agg[k][0] += n; agg[k][1] += s; agg[k][2] += t
    tot[0] += n; tot[1] += s
print('total warp-inst %d samples %d' % tuple(tot))
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][SORTKEY])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    txt = ''
    if k:
        f = k[0]
        if f not in src:
            try: src[f] = open(os.environ.get("SRC_DIR", "cdftools_b200/csrc/") + f).read().split('\n')
            except Exception: src[f] = []
        if 0 < k[1] <= len(src[f]): txt = src[f][k[1] - 1].strip()[:80]
    print('%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  %s:%s  %s' % (100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1), v[2] / max(v[0], 1), k[0] if k else '?', k[1] if k else '', txt))
