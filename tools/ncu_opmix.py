"""Aggregate an `ncu --page source --csv` dump by opcode and list the hottest SASS lines.
usage: python tools/ncu_opmix.py file.csv [ntop]"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ie, src, ns = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
ti = hdr.index('Thread Instructions Executed')
cnt = collections.Counter(); tot = 0; samp = collections.Counter(); stot = 0
lines = []
for r in rows[2:]:
    try: n = int(r[ie]); s = int(r[ns]); t = int(r[ti])
    except Exception: continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[src])
    op = m.group(2).split('.')[0] if m else '?'
    cnt[op] += n; tot += n; samp[op] += s; stot += s
    lines.append((s, n, t, r[src].strip()))
print('total warp-instructions', tot, 'samples', stot)
for k, v in cnt.most_common(28): print('  %-12s %12d %5.1f%%   samples %5.1f%%' % (k, v, 100 * v / tot, 100 * samp[k] / max(stot, 1)))
print('hottest lines by samples:')
for s, n, t, txt in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print('  %6d smp %10d exec  thr/inst %4.1f  %s' % (s, n, t / max(n, 1), txt[:90]))
