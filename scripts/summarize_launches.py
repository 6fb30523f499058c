"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print('%-44s %7s %12s %10s %7s' % ('kernel', 'n', 'total_us', 'avg_us', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-44s %7d %12.1f %10.2f %6.1f%%' % (k[:44], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print('%-44s %7d %12.1f' % ('TOTAL', sum(v[0] for v in agg.values()), tot))


if __name__ == '__main__':
    main(sys.argv[1])
