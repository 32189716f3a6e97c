"""DRAM traffic per launch of one kernel family from an ncu CSV
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:<family> --csv`), merged
into profiles/traffic.json under the bench config it was captured on -- bench.py reports it as roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/traffic_c2.csv c2 conv_tc_kernel [profiles/traffic.json]
"""
import collections
import csv
import json
import os
import sys

SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    path, cfg, family = sys.argv[1:4]
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                             'profiles', 'traffic.json')
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    per = collections.defaultdict(dict)
    for row in csv.DictReader(lines):
        if family not in row['Kernel Name']:
            continue
        v = float(row['Metric Value'].replace(',', ''))
        per[int(row['ID'])][row['Metric Name']] = v * SCALE.get(row['Metric Unit'], 1.0)
    ids = sorted(per)
    ids = ids[len(ids) // 2:]          # the warm-up step comes first; keep the timed step
    tot = sum(per[i].get('dram__bytes_read.sum', 0.0) + per[i].get('dram__bytes_write.sum', 0.0) for i in ids)
    entry = {'bytes_per_launch': tot / max(1, len(ids)), 'launches': len(ids),
             'source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d %s launches of one step (%s)'
                       % (len(ids), family, os.path.basename(path))}
    data = {}
    if os.path.exists(out):
        with open(out) as f:
            data = json.load(f)
    data.setdefault(cfg, {})[family] = entry
    with open(out, 'w') as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print(cfg, family, entry)


if __name__ == '__main__':
    main()
