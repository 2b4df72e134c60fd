"""Timing experiments: per-kernel device time of days 85..95 (epidemic peak) with one sweep stage skipped.
Results are NOT valid simulations; this only attributes the sweep's time to its stages."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for flag in (0, 1, 2, 3):
    ctx = bench.make_context(bench.workload_spec('hus'), R, 0, 180, seed=1)
    ctx.run(85)
    while len(ctx._plan) < 180: ctx._plan_next_day()
    ctx._engine.set_schedule(0, ctx._plan[:180])
    lib = ctx._engine.lib.dll
    lib.rb_debug_flag.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.rb_debug_flag(ctx._engine.h, flag)
    k = ctx._engine.step_profiled(10)
    print('skip stage', flag, {n: round(float(v) / 10 * 1000, 1) for n, v in zip(['pre', 'sweep', 'expose', 'resolve', 'post'], k)}, flush=True)
    ctx.close()
