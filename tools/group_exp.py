"""Replica-group experiment (measurement aid): device time of the HUS ensemble step for several (groups, wave %) settings,
and a check that every setting gives bit-identical stats rows.
    python tools/group_exp.py --replicas 256 --configs 1:100,2:50,2:75,2:100,3:34,4:25,4:50"""
import argparse
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--replicas', type=int, default=256)
ap.add_argument('--days', type=int, default=180)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--configs', default='1:100,2:50,2:75,2:100,3:34,4:25,4:50')
a = ap.parse_args()
ref = None
for cfg in a.configs.split(','):
    ng, pct = cfg.split(':')[:2]
    os.environ.pop('RB_L2_FETCH', None)
    if len(cfg.split(':')) > 2:
        os.environ['RB_L2_FETCH'] = cfg.split(':')[2]
    os.environ['RB_GROUPS'] = ng
    os.environ['RB_GROUP_WAVE_PCT'] = pct
    ctx = bench.make_context(bench.workload_spec('hus'), a.replicas, 0, a.days, seed=1)
    ms = []
    for step in range(2 + a.steps):
        ctx.reset(1000 + step)
        ctx.run(a.days)
        if step >= 2:
            ms.append(ctx._engine.last_step_ms())
    rows = ctx.series(0, a.days)
    h = hashlib.sha1(np.ascontiguousarray(rows).tobytes()).hexdigest()[:12]
    if ref is None:
        ref = h
    ctx.close()
    v = bench.N_HUS * a.days * a.replicas / (np.mean(ms) / 1e3)
    print('groups %s wave %s%% %s: %.2f ms/step (min %.2f)  %.3e agent-days/s  rows %s %s'
          % (ng, pct, cfg, np.mean(ms), np.min(ms), v, h, 'OK' if h == ref else 'DIFFER'), flush=True)
