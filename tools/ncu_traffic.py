"""DRAM traffic per launch of every per-day kernel from an ncu CSV, stamped with the kernel sources it was measured on:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file traffic.csv python tools/prof_run.py --replicas 256 --days 180
    python tools/ncu_traffic.py traffic.csv 256 180 profiles/r02_dram_traffic_R256.json "<the command>"
bench.py reads the JSON for roofline.traffic and prints null instead when the stamp no longer matches the sources."""
import csv
import json
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    path, R, days, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    command = sys.argv[5] if len(sys.argv) > 5 else None
    lines = [l for l in open(path) if l.startswith('"')]
    per = defaultdict(dict)
    name = {}
    for r in csv.DictReader(lines):
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        if r['Metric Name'].startswith('dram__bytes'):
            v *= dict(byte=1, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9)[u]
        elif r['Metric Name'] == 'gpu__time_duration.sum':
            v = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)      # -> us
        per[r['ID']][r['Metric Name']] = v
        name[r['ID']] = r['Kernel Name'].split('(')[0].split('<')[0]
    agg = defaultdict(lambda: dict(launches=0, read=0.0, write=0.0, us=0.0))
    for i, m in per.items():
        a = agg[name[i]]
        a['launches'] += 1
        a['read'] += m.get('dram__bytes_read.sum', 0.0)
        a['write'] += m.get('dram__bytes_write.sum', 0.0)
        a['us'] += m.get('gpu__time_duration.sum', 0.0)
    kernels = {k: dict(launches=a['launches'], mean_read_bytes=a['read'] / a['launches'], mean_write_bytes=a['write'] / a['launches'],
                       mean_traffic_bytes=(a['read'] + a['write']) / a['launches'], total_us_under_ncu=a['us'])
               for k, a in agg.items()}
    res = dict(source_stamp=bench.kernel_source_stamp(), replicas=R, days=days, command=command,
               note='per-launch means over every launch of one run; launch = one replica group x one day', kernels=kernels)
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
