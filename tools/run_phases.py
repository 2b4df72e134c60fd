"""Where a persistent-kernel day goes (measurement aid): ns per phase on the lead CTA, barrier waits included.
    RB_PERSISTENT=1 RB_RUN_CTAS=148 python tools/run_phases.py [R]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = bench.make_context(bench.workload_spec('hus'), R, 0, 180, seed=1)
f = ctx._engine.lib.f
ctx.run(180)
ctx.reset(5)
f['debug_flag'](ctx._engine.h, 8)
ctx.run(180)
out = np.zeros(16, dtype=np.int64)
f['debug_phase_cycles'](ctx._engine.h, 0, out.ctypes.data)
names = ['first pre', 'sweep (own work)', 'sweep barrier wait', 'expose + barrier', 'resolve + barrier', 'boundary + barrier']
print('R=%d RUN_CTAS=%s step %.3f ms' % (R, os.environ.get('RB_RUN_CTAS'), ctx._engine.last_step_ms()))
for n, v in zip(names, out[8:14]):
    print('  %-22s %8.1f us/day' % (n, v / 180 / 1e3))
