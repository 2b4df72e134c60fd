"""Instruction-mnemonic histogram of every kernel of libreina_b200.so (cuobjdump -sass), for profiles/: what the ISA-level
claims rest on (LDGSTS / UBLKCP / UTMA* = asynchronous copies, ATOM / RED = atomics, SHFL / VOTE = warp collectives ...).
    python tools/sass_histogram.py [lib.so] > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'reina_b200', 'libreina_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and cur is not None:
        cur[m.group(1)] += 1
arch = re.search(r'arch = (sm_\w+)', txt)
print('SASS instruction histogram of %s (%s), cuobjdump -sass' % (os.path.basename(lib), arch.group(1) if arch else '?'))
for fn, c in per.items():
    name = subprocess.run(['c++filt', fn], stdout=subprocess.PIPE, text=True).stdout.strip()
    tot = sum(c.values())
    print('\n%s: %d instructions' % (name, tot))
    print('  ' + '  '.join('%s %d' % kv for kv in c.most_common(28)))
