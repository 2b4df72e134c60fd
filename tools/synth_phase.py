"""Cycle breakdown of the day-boundary kernel per phase on the synthetic large population:
    python tools/synth_phase.py [N] [days] [stretch]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth_run  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
days = int(sys.argv[2]) if len(sys.argv) > 2 else 180
stretch = int(sys.argv[3]) if len(sys.argv) > 3 else 20
ctx = synth_run.make(n, days)
lib = ctx._engine.lib.dll
lib.rb_debug_flag.argtypes = [ctypes.c_void_p, ctypes.c_int32]
lib.rb_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
lib.rb_debug_flag(ctx._engine.h, 9)
names = {0: 'stats row', 1: 'imports+init_day', 10: 'queue sort', 2: 'queue drain', 5: 'trace l0', 6: 'trace l1+edges', 7: 'trace cycles', 3: 'trace finish', 4: 'vaccination', 8: 'event sort', 9: 'capacity scan'}
prev = np.zeros(16, dtype=np.int64)
G = len(ctx.age_group_labels)
for lo in range(0, days, stretch):
    m = min(stretch, days - lo)
    ctx.run(m)
    out = np.zeros(16, dtype=np.int64)
    lib.rb_debug_phase_cycles(ctx._engine.h, 0, out.ctypes.data)
    d = out - prev
    prev = out
    rows = ctx.series(lo, m)[0]
    sc = rows[:, 13 * G:]
    print('days %3d-%3d us/day:' % (lo, lo + m), {names[k]: round(float(d[k]) / m / 1900.0, 1) for k in names},
          'run ms %.2f' % ctx._engine.last_step_ms(), 'max queue %d max infected %d' % (sc[:, 7].max(), rows[:, 2 * G:3 * G].sum(axis=1).max()), flush=True)
