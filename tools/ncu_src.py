"""Per-source-line instruction / stall-sample shares, per kernel, from
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_src.py x.csv [top] [kernel-substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else ''
per = {}
hdr = None
fpath = func = ''
for r in rows:
    if r and r[0] == 'File Path':
        fpath = r[1]
    elif r and r[0] == 'Function Name':
        func = r[1]
    elif r and r[0] == 'Line No':
        hdr = r
        iS, iI, iT = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    elif hdr and len(r) > iT and r[0].isdigit():
        try:
            per.setdefault(func, []).append((fpath.split('/')[-1], int(r[0]), r[1], int(r[iS]), int(r[iI]), int(r[iT])))
        except ValueError:
            pass
for func, out in per.items():
    if only not in func:
        continue
    ts = sum(d[3] for d in out) or 1
    ti = sum(d[4] for d in out) or 1
    print('=== %s: %d samples, %d warp instructions' % (func, ts, ti))
    out.sort(key=lambda d: -(d[3] / ts + d[4] / ti))
    for d in out[:top]:
        print('%-12s %5d inst %5.1f%% smp %5.1f%% eff %4.1f | %s' % (d[0][:12], d[1], 100 * d[4] / ti, 100 * d[3] / ts, d[5] / max(d[4], 1), d[2].strip()[:105]))
