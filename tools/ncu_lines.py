"""Join an ncu report's SASS-level samples with the source lines of the profiled cubin.

    python tools/ncu_lines.py <report.ncu-rep> <lib.so used for the capture> <kernel name> [top N]

ncu's CSV export of the source page carries per-SASS-instruction metrics but no line numbers; nvdisasm -g gives the
line of every SASS instruction of the same cubin in the same order.  Prints the source lines with the most stall
samples and the most executed instructions."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], check=True, stdout=subprocess.PIPE, text=True).stdout
    out, cur, on = [], None, False
    for line in txt.splitlines():
        m = re.match(r'\s*//-+ \.text\.(\S+)', line)
        if m:
            on = kernel in m.group(1)
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
            out.append(cur)
    return out


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + kernel], check=True,
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    data = [r for r in rows[hdr_i + 1:] if r and r[0].startswith('0x')]
    lines = sass_lines(lib, kernel)
    if len(lines) != len(data):
        print('warning: %d SASS instructions in the cubin vs %d in the report' % (len(lines), len(data)))
    si, ii, ti = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    agg = collections.defaultdict(lambda: [0, 0, 0])
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    stalls = collections.defaultdict(lambda: collections.Counter())
    for k, r in enumerate(data):
        key = lines[k] if k < len(lines) else None
        agg[key][0] += int(r[si] or 0); agg[key][1] += int(r[ii] or 0); agg[key][2] += int(r[ti] or 0)
        for c in stall_cols:
            if r[c] and r[c] != '0':
                stalls[key][hdr[c]] += int(r[c])
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    src = {}
    print('kernel %s: %d samples, %d warp instructions' % (kernel, tot_s, tot_i))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ''
        if key:
            path = next((p for p in (os.path.join('reina_b200/csrc', key[0]), key[0]) if os.path.exists(p)), None)
            if path:
                if path not in src:
                    src[path] = open(path).read().splitlines()
                text = src[path][key[1] - 1].strip()[:90]
        st = ', '.join('%s %d' % (n.replace('stall_', ''), c) for n, c in stalls[key].most_common(3))
        print('%-16s smp %5.1f%% inst %5.1f%% eff %4.1f | %s  [%s]' % (
            '%s:%d' % key if key else '?', 100.0 * v[0] / tot_s, 100.0 * v[1] / tot_i, v[2] / max(v[1], 1), text, st))


if __name__ == '__main__':
    main()
