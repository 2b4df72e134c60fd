"""Cycle breakdown of the day-boundary kernel (k_between) per phase, for one seed: python tools/phase_run.py [R]"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = bench.make_context(bench.workload_spec('hus'), R, 0, 180, seed=1)
lib = ctx._engine.lib.dll
lib.rb_debug_flag.argtypes = [ctypes.c_void_p, ctypes.c_int32]
lib.rb_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
lib.rb_debug_flag(ctx._engine.h, 9)
names = {0: 'stats row', 1: 'imports+init_day', 2: 'queue drain', 3: 'contact tracing', 4: 'vaccination', 8: 'event sort', 9: 'capacity scan'}
prev = np.zeros(16, dtype=np.int64)
for lo, hi in ((0, 60), (60, 118), (118, 180)):
    ctx.run(hi - lo)
    out = np.zeros(16, dtype=np.int64)
    lib.rb_debug_phase_cycles(ctx._engine.h, 0, out.ctypes.data)
    d = out - prev; prev = out
    print('days %d-%d (us/day at 1.9 GHz):' % (lo, hi), {names[k]: round(float(d[k]) / (hi - lo) / 1900.0, 1) for k in names}, 'run ms %.2f' % ctx._engine.last_step_ms(), flush=True)
