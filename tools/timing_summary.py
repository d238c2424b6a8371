"""Summarise tools/layer_timing.py output: per-category totals and the slowest launches."""
import collections
import re
import sys

rows = []
for l in open(sys.argv[1]):
    m = re.match(r'\[pdes timing\]\s+([\d.]+) us\s+(.*)', l)
    if m:
        rows.append((float(m.group(1)), m.group(2)))
tot = sum(r[0] for r in rows)
print('total %.1f us over %d launches' % (tot, len(rows)))
agg = collections.OrderedDict()
for t, n in rows:
    k = n.split()[0]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += t
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-20s n=%3d %8.1f us %5.1f%%' % (k, n, t, 100 * t / tot))
if len(sys.argv) > 2:
    for t, n in rows:
        print('%8.1f  %s' % (t, n))
else:
    for t, n in sorted(rows, reverse=True)[:30]:
        print('%8.1f  %s' % (t, n))
