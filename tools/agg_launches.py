"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    agg_launches.py launches.csv [delimiter-substring] [--list]

With a delimiter (e.g. p_sample_step, adamw) only the launches after the second-to-last and up to the last kernel whose
name contains it are aggregated: exactly one step of a loop, whatever the launch count per step is."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
delim = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mv, mu, gs = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
data = []
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r'\(.*', '', r[kn])
    v = float(r[mv].replace(',', ''))
    v = v / 1000 if r[mu] == 'ns' else (v * 1000 if r[mu] == 'ms' else v)
    data.append((name, v, r[gs]))
if delim:
    idx = [i for i, d in enumerate(data) if delim in d[0]]
    if len(idx) < 2:
        raise SystemExit(f"need two '{delim}' launches to delimit a step, found {len(idx)}")
    data = data[idx[-2] + 1:idx[-1] + 1]
if "--list" in sys.argv:
    for i, (n, v, g) in enumerate(data):
        print(i, f"{v:8.1f}", g, n.replace('void fcwdm::', '').replace('fcwdm::', '')[:60])
    sys.exit(0)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for name, v, _ in data:
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"{'us':>10} {'n':>5} {'share':>6}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {n:5d} {100 * t / tot:5.1f}%  {k[:110]}")
print(f"{tot:10.1f} {sum(v[0] for v in agg.values()):5d} 100.0%  TOTAL (serialised, cold-cache launch durations)")
