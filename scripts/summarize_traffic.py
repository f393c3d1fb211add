"""Per-kernel DRAM traffic / duration / occupancy from an ncu --csv run with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... : python scripts/summarize_traffic.py traffic.csv [out.json]"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mn, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    per.setdefault(r[idc], {'k': re.sub(r'\(.*', '', r[kn])[:80]})[r[mn]] = float(r[mv].replace(',', ''))
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d['k'], collections.Counter())
    a['n'] += 1
    for k, v in d.items():
        if k != 'k':
            a[k] += v
out = {}
print('%5s %9s %10s %10s %8s %6s  kernel' % ('n', 'avg_us', 'rd_MB', 'wr_MB', 'GB/s', 'occ%'))
for k, a in sorted(agg.items(), key=lambda x: -x[1]['gpu__time_duration.sum']):
    n = a['n']
    t = a['gpu__time_duration.sum'] / n / 1e3
    rd, wr = a['dram__bytes_read.sum'] / n, a['dram__bytes_write.sum'] / n
    occ = a['sm__warps_active.avg.pct_of_peak_sustained_active'] / n
    print('%5d %9.2f %10.3f %10.3f %8.1f %6.1f  %s' % (n, t, rd / 1e6, wr / 1e6, (rd + wr) / t / 1e3 if t else 0, occ, k))
    out[k] = {'launches': n, 'avg_us': round(t, 2), 'dram_read_bytes_per_launch': int(rd), 'dram_write_bytes_per_launch': int(wr),
              'warps_active_pct': round(occ, 1)}
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
