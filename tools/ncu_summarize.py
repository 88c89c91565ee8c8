#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
tot = defaultdict(lambda: [0.0, 0])
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'<.*', lambda m: m.group(0)[:60], name)
    unit = row.get('Metric Unit', 'ns')
    v = float(row['Metric Value'].replace(',', ''))
    scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}.get(unit, 1e-3)
    tot[name][0] += v * scale
    tot[name][1] += 1
allt = sum(v[0] for v in tot.values())
print(f'total {allt/1e3:.3f} ms over {sum(v[1] for v in tot.values())} launches')
print(f'{"share":>7} {"ms":>10} {"count":>7} {"us/launch":>10}  kernel')
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f'{100*us/allt:6.2f}% {us/1e3:10.3f} {n:7d} {us/n:10.2f}  {name[:110]}')
