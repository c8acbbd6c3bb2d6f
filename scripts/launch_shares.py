#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel count, average, median, share."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = defaultdict(list)
for r in rows[start:]:
    if len(r) > vi:
        try:
            d[r[ki][:70]].append(float(r[vi].replace(',', '')))
        except ValueError:
            pass
tot = sum(sum(v) for v in d.values())
print('ncu launch list %s: per-launch times are cold-cache and serialised -- compare SHARES' % sys.argv[1])
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    print('%-72s n=%4d avg=%8.2f us med=%8.2f us share=%5.1f%%' % (k, len(v), sum(v) / len(v) / 1000, v2[len(v2) // 2] / 1000, 100 * sum(v) / tot))
