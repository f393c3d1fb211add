"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg, total = collections.OrderedDict(), 0.0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    t = float(r[mv].replace(',', ''))
    short = re.sub(r'\(.*', '', r[kn])[:100]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += t
    total += t
print('launches %d  total %.1f us' % (sum(a[0] for a in agg.values()), total / 1e3))
print('%7s %12s %8s %7s  %s' % ('count', 'total_us', 'avg_us', 'share', 'kernel'))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print('%7d %12.1f %8.2f %6.1f%%  %s' % (n, t / 1e3, t / 1e3 / n, 100 * t / total, k))
