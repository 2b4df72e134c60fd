"""Synthetic large population on one GPU (BASELINE configs[4] population, unsharded): HUS age histogram scaled to
N agents, beds / ICU / imports scaled alike.  python tools/synth_run.py [N] [days]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reina_b200 import inputs, model  # noqa: E402


def make(n_agents, days, seed=0, n_replicas=1):
    f = n_agents / 1685983.0
    v = inputs.default_variables()
    v['hospital_beds'], v['icu_units'] = int(round(2600 * f)), int(round(300 * f))
    ivs = []
    for iv in v['interventions']:
        iv = list(iv)
        if iv[0] in ('import-infections', 'import-infections-weekly'):
            iv[2] = int(round(iv[2] * f))
        ivs.append(iv)
    v['interventions'] = ivs
    counts = inputs.synthetic_age_counts(n_agents)
    args = inputs.build_context_args(v, age_count_override=counts)
    args['random_seed'] = seed
    ctx = model.Context(n_replicas=n_replicas, max_days=days + 1, **args)
    for iv in inputs.active_interventions(v):
        ctx.add_intervention(iv)
    return ctx


if __name__ == '__main__':
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
    days = int(sys.argv[2]) if len(sys.argv) > 2 else 180
    t0 = time.time()
    ctx = make(n, days)
    print('create %.1f s' % (time.time() - t0), flush=True)
    ctx.run(days)
    ms = ctx._engine.last_step_ms()
    rows = ctx.series(0, days)[0]
    G = len(ctx.age_group_labels)
    tot = lambda a: rows[:, a * G:(a + 1) * G].sum(axis=1)
    print('N=%d days=%d device %.1f ms -> %.3e agent-days/s; all_infected %d dead %d peak infected %d' % (
        n, days, ms, n * days / (ms / 1e3), tot(3)[-1], tot(9)[-1], tot(2).max()))
    assert (tot(0) + tot(2) + tot(10) + tot(9) == n).all()
