"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by (file, line).
usage: python scripts/ncu_hot_lines.py export.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
agg = collections.defaultdict(lambda: [0, 0])
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        ci, cs = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if cur is None or hdr is None or r[0] == '':
        continue
    try:
        ln = int(r[0]); n = int(r[ci] or 0); s = int(r[cs] or 0)
    except ValueError:
        continue
    agg[(cur, ln)][0] += n
    agg[(cur, ln)][1] += s
    src[(cur, ln)] = r[1]
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
print('total warp-instructions %d, samples %d' % (tot, tots))
for (f, ln), (n, s) in sorted(agg.items(), key=lambda x: -x[1][int(len(sys.argv) > 3)])[:top]:
    print('%-16s %5d  inst %5.1f%%  samples %5.1f%%  %s' % (f[:16], ln, 100.0 * n / max(tot, 1), 100.0 * s / max(tots, 1), src[(f, ln)].strip()[:100]))
