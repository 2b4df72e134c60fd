"""Instruction / stall-sample shares of k_sweep per code region, from the source page CSV of an ncu report:
    python tools/ncu_regions.py X.ncu-rep [kernel-substring]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else 'k_sweep'
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
REGIONS = [('sweep.cuh', 40, 45, 'ring_push'), ('sweep.cuh', 69, 118, 'stage E (contact count, items)'),
           ('sweep.cuh', 119, 126, 'emit_event'), ('sweep.cuh', 127, 225, 'stage T (transitions)'),
           ('sweep.cuh', 226, 256, 'stage 1 (active lane)'), ('sweep.cuh', 257, 290, 'cp.async / push4'),
           ('sweep.cuh', 281, 345, 'kernel prologue'), ('sweep.cuh', 346, 421, 'producer'), ('sweep.cuh', 422, 470, 'consumer glue'),
           ('rng.cuh', 0, 9999, 'rng.cuh (philox, feistel, gamma)'), ('state.cuh', 0, 9999, 'state.cuh (age_of, sweep_pos, ...)')]
hdr = None
fpath = func = ''
agg = {}
tot = [0, 0]
for r in rows:
    if r and r[0] == 'File Path':
        fpath = r[1].split('/')[-1]
    elif r and r[0] == 'Function Name':
        func = r[1]
    elif r and r[0] == 'Line No':
        hdr = r
        iS, iI = hdr.index('# Samples'), hdr.index('Instructions Executed')
    elif hdr and len(r) > iI and r[0].isdigit() and only in func:
        try:
            ln, s, i = int(r[0]), int(r[iS]), int(r[iI])
        except ValueError:
            continue
        name = 'other: ' + fpath
        for f, lo, hi, nm in REGIONS:
            if fpath == f and lo <= ln <= hi:
                name = nm
                break
        a = agg.setdefault(name, [0, 0])
        a[0] += s; a[1] += i
        tot[0] += s; tot[1] += i
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-42s inst %5.1f%%  samples %5.1f%%' % (k, 100.0 * i / max(tot[1], 1), 100.0 * s / max(tot[0], 1)))
