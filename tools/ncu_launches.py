"""Summaries of the ncu launch-list CSVs written by tools/collect_profiles.sh:
    python tools/ncu_launches.py gpurun_out/prof/launches_R128.csv gpurun_out/prof/sweep_traffic_R128.csv out.json"""
import csv
import json
import sys
from collections import defaultdict


def rows_of(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))


def main():
    launches, traffic, out = sys.argv[1:4]
    dur = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows_of(launches):
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        name = r['Kernel Name'].split('(')[0]
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        us = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
        dur[name] += us
        cnt[name] += 1
    tot = sum(dur.values())
    res = dict(launch_list=dict(total_us=tot, kernels={k: dict(launches=cnt[k], us=dur[k], share=dur[k] / tot) for k in sorted(dur, key=dur.get, reverse=True)}))
    per = defaultdict(dict)
    for r in rows_of(traffic):
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        if r['Metric Name'].startswith('dram__bytes'):
            v *= dict(byte=1, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9)[u]
        per[r['ID']][r['Metric Name']] = v
    n = len(per)
    rd = sum(p['dram__bytes_read.sum'] for p in per.values()) / n
    wr = sum(p['dram__bytes_write.sum'] for p in per.values()) / n
    res['k_sweep_dram'] = dict(launches=n, mean_read_bytes=rd, mean_write_bytes=wr, mean_traffic_bytes=rd + wr)
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
