"""Per-kernel device time (us/day) of the HUS ensemble in three stretches of the epidemic:
    [REINA_B200_LIB=path/to/variant.so] python tools/kern_times.py [R]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = bench.make_context(bench.workload_spec('hus'), R, 0, 180, seed=1)
while len(ctx._plan) < 180:
    ctx._plan_next_day()
ctx._engine.set_schedule(0, ctx._plan[:180])
names = ['pre', 'sweep', 'expose', 'resolve', 'post']
tot = [0.0] * 5
line = []
for lo, hi in ((0, 60), (60, 120), (120, 180)):
    k = ctx._engine.step_profiled(hi - lo)
    for i in range(5):
        tot[i] += float(k[i])
    line.append('d%d-%d ' % (lo, hi) + ' '.join('%s %.0f' % (n, float(v) / (hi - lo) * 1000) for n, v in zip(names, k)))
print(os.environ.get('REINA_B200_LIB', 'default'), '|', ' | '.join(line), '| total ms %.1f sweep ms %.1f' % (sum(tot), tot[1]), flush=True)
