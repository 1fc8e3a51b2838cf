"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST step.
usage: python tools/summarize_launches.py launches.csv [launches_in_step]"""
import csv
import re
import sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
if n:
    rows = rows[-n:]


def name(r):
    k = re.sub(r'\(.*', '', r['Kernel Name'])
    return k.replace('void ', '').replace('rcfd::<unnamed>::', '')


tot = sum(float(r['Metric Value']) for r in rows) / 1e6
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    a = agg[name(r)]
    a[0] += 1
    a[1] += float(r['Metric Value']) / 1e6
print('launches %d total ms %.3f' % (len(rows), tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-70s %5d %9.3f ms %5.1f%%' % (k[:70], c, t, 100 * t / tot))
