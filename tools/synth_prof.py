"""Per-kernel device time of the synthetic large population on one GPU, in stretches of days:
    python tools/synth_prof.py [N] [days] [stretch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth_run  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
days = int(sys.argv[2]) if len(sys.argv) > 2 else 180
stretch = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ctx = synth_run.make(n, days)
while len(ctx._plan) < days:
    ctx._plan_next_day()
ctx.upload_inputs()
ctx._engine.set_schedule(0, ctx._plan[:days])
names = ['pre', 'sweep', 'expose', 'resolve', 'post']
tot = 0.0
for lo in range(0, days, stretch):
    m = min(stretch, days - lo)
    k = ctx._engine.step_profiled(m)
    tot += float(k.sum())
    print('days %3d-%3d us/day:' % (lo, lo + m), {a: round(float(v) / m * 1000, 1) for a, v in zip(names, k)}, flush=True)
print('N=%d: %.1f ms for %d days -> %.3e agent-days/s (profiled launches)' % (n, tot, days, n * days / (tot / 1e3)))
