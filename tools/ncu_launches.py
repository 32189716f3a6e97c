"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total time, share and count.
   python tools/ncu_launches.py gpurun_out/launches.csv [first_id last_id]"""
import collections
import csv
import re
import sys


def short(name):
    name = name.replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    name = re.sub(r'^void\s+', '', name)
    m = re.match(r'([A-Za-z0-9_:]+)', name)
    return m.group(1) if m else name[:60]


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum' or not (lo <= int(row['ID']) <= hi):
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}[row['Metric Unit']]
        k = short(row['Kernel Name'])
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    print('total %.2f ms over %d launches (ids %d..%d)' % (T, sum(cnt.values()), lo, min(hi, 10 ** 9)))
    print('%10s %7s %6s  kernel' % ('ms', 'share', 'count'))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:30]:
        print('%10.3f %6.1f%% %6d  %s' % (v, 100 * v / T, cnt[k], k))


if __name__ == '__main__':
    main()
