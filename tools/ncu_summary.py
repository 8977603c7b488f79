"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        key = row['Kernel Name'].split('(')[0].replace('void ', '').replace('nats::<unnamed>::', '')
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print('# %s: %d launches, %.1f us of kernel time (cold-cache, serialised: compare SHARES)' % (path, n, tot))
    print('%10s %7s %7s %10s  %s' % ('total_us', 'share', 'count', 'avg_us', 'kernel'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%10.1f %6.1f%% %7d %10.2f  %s' % (v[1], 100 * v[1] / tot, v[0], v[1] / v[0], k))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
