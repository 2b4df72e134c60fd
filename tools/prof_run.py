"""Short run of the hot path for ncu captures: python tools/prof_run.py --replicas 32 --days 100 [--workload hus]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--replicas', type=int, default=32)
ap.add_argument('--days', type=int, default=100)
ap.add_argument('--workload', default='hus')
a = ap.parse_args()
ctx = bench.make_context(bench.workload_spec(a.workload), a.replicas, 0, a.days, seed=1)
ctx.run(a.days)
rows = ctx.series(0, a.days)
G = len(ctx.age_group_labels)
print('infected on last day (replica 0):', rows[0, -1, 2 * G:3 * G].sum())
