"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r'\(.*', '', r[kn])
    v = float(r[mv].replace(',', ''))
    v = v / 1000 if r[mu] == 'ns' else (v * 1000 if r[mu] == 'ms' else v)
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"{'us':>10} {'n':>5} {'share':>6}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {n:5d} {100 * t / tot:5.1f}%  {k[:110]}")
print(f"{tot:10.1f} {sum(v[0] for v in agg.values()):5d} 100.0%  TOTAL (serialised, cold-cache launch durations)")
