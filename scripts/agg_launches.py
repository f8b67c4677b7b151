"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python scripts/agg_launches.py in.csv [out.csv]"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(row['Metric Unit'], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out = [('kernel', 'launches', 'total_ms', 'avg_us', 'share_pct')]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append((k, v[0], round(v[1] / 1e3, 3), round(v[1] / v[0], 1), round(100 * v[1] / tot, 2)))
out.append(('TOTAL', sum(v[0] for v in agg.values()), round(tot / 1e3, 3), '', 100.0))
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w', newline='') as fh:
        csv.writer(fh).writerows(out)
for r in out[:24]:
    print(*r, sep='\t')
