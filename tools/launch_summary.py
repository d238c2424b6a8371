"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of
the LAST profiled training step and the slowest individual launches."""
import collections
import csv
import re
import sys


def main(path, steps=2, top=25):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((int(row['ID']), re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('unnamed>::', ''),
                         float(row['Metric Value'].replace(',', '')), row['Grid Size'], row['Block Size']))
    per = len(rows) // steps
    step = rows[-per:]
    agg = collections.OrderedDict()
    for _, k, v, g, b in step:
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    print('%d launches in the last step, %.1f us total (cold-cache, serialised: compare shares)' % (len(step), tot / 1000))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-50s n=%3d  %9.1f us  %5.1f%%' % (k[:50], n, v / 1000, 100 * v / tot))
    print()
    for i, k, v, g, b in sorted(step, key=lambda r: -r[2])[:top]:
        print('%5d %-42s %8.1f us  grid %-14s block %s' % (i, k[:42], v / 1000, g, b))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
