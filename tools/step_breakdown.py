"""Per-kernel totals of ONE training step from an ncu launch list (the step between the last two adam_kernel launches).
usage: python tools/step_breakdown.py launches.csv [--timeline]"""
import csv
import re
import sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if 'adam_kernel' in r['Kernel Name']]
step = rows[idx[-2] + 1:idx[-1] + 1] if len(idx) >= 2 else rows


def nm(r):
    return re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('rcfd::<unnamed>::', '')


tot = sum(float(r['Metric Value']) for r in step) / 1e6
print('launches in step %d, total %.3f ms (serialised, cold caches)' % (len(step), tot))
if '--timeline' in sys.argv:
    for i, r in enumerate(step):
        print(i, nm(r)[:44], r['Grid Size'], '%.1f' % (float(r['Metric Value']) / 1e3))
    sys.exit(0)
agg = defaultdict(lambda: [0, 0.0])
for r in step:
    a = agg[nm(r)]
    a[0] += 1
    a[1] += float(r['Metric Value']) / 1e6
groups = defaultdict(float)
for k, (c, t) in agg.items():
    g = 'conv fwd/dgrad' if k.startswith('conv_') else 'wgrad' if k.startswith('wgrad') else 'batch norm' if k.startswith('bn_') \
        else 'pack/unpack' if 'pack' in k else 'other'
    groups[g] += t
for g, t in sorted(groups.items(), key=lambda kv: -kv[1]):
    print('  %-16s %7.3f ms %5.1f%%' % (g, t, 100 * t / tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-64s %5d %9.3f ms %5.1f%%' % (k[:64], c, t, 100 * t / tot))
