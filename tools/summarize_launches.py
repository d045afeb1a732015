"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<round>_launches_summary.txt"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1e3 if row['Metric Unit'] == 'ns' else (v * 1e3 if row['Metric Unit'] == 'ms' else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f'# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms of kernel time (ncu: cold-cache, serialised; compare SHARES)')
    print(f'{"total ms":>10} {"share":>6} {"n":>5} {"avg us":>10}  kernel')
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{us / 1e3:10.3f} {100 * us / tot:5.1f}% {n:5d} {us / n:10.1f}  {k[:110]}')


if __name__ == '__main__':
    main(sys.argv[1])
