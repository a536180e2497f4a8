#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (second half of the launches = the
last of two identical steps)."""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
data = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[1:]]
half = data[len(data) // 2:]
t = defaultdict(lambda: [0, 0.0])
for k, v in half:
    k = k.replace('void ', '').replace('at::native::', '').replace('<unnamed>::', '')
    k = re.sub(r'\(.*', '', k)[:100]
    t[k][0] += 1; t[k][1] += v
tot = sum(v[1] for v in t.values())
print(f'{len(half)} launches, {tot / 1e3:.1f} us of kernel time')
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f'{v[0]:4d} {v[1] / 1e3:8.1f} us {v[1] / tot * 100:5.1f}%  {k}')
