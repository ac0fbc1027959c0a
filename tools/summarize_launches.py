#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*.csv)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    n = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('dsb200::', '')
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1000 if row['Metric Unit'] == 'ns' else v * 1000 if row['Metric Unit'] == 'ms' else v
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print('{:>10s} {:>5s} {:>6s}  kernel'.format('us', 'n', 'share'))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('{:10.1f} {:5d} {:5.1f}%  {}'.format(t, c, 100 * t / tot, n))
print('{:10.1f} {:5d}         total'.format(tot, sum(c for c, _ in agg.values())))
