"""Drive the UNMODIFIED reference engine (oracle/_ref, built by oracle/build_ref.sh).

TEST INFRASTRUCTURE ONLY -- imported by tests/, tests/golden/make_golden.py and bench.py's
reference / cpu_baseline legs, never by anything under reina_b200/.

Restates the caller side of the hot path, calc/simulation.py:151-290: build population /
healthcare / disease parameters, `model.Context(...)`, `add_intervention`, then per day
`generate_state()` followed by `iterate()`; only `iterate()` is timed (what SURVEY.md section 6 reports).
Generated mobility / vaccination interventions (common/interventions.py:358-364) need files that
are downloaded from the network and are defined as empty for every benchmark config.
"""
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

POP_SERIES = ['susceptible', 'vaccinated', 'infected', 'all_infected', 'detected', 'all_detected',
              'in_icu', 'cum_icu', 'in_ward', 'dead', 'recovered', 'non_hospital_deaths',
              'new_infections']
SCALAR_SERIES = ['available_icu_units', 'available_hospital_beds', 'total_icu_units', 'r',
                 'exposed_per_day', 'ct_cases_per_day', 'mobility_limitation']
PLACES = ['home', 'work', 'school', 'transport', 'leisure', 'other']


def available():
    import sysconfig
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    return os.path.exists(os.path.join(_HERE, '_ref', 'cythonsim', 'main' + ext))


def load_model():
    """Import oracle/_ref/cythonsim/main through the three import stubs."""
    if not available():
        raise RuntimeError('oracle/_ref is not built; run oracle/build_ref.sh where /root/reference exists')
    for p in (os.path.join(_HERE, '_ref'), os.path.join(_HERE, 'stubs')):
        if p not in sys.path:
            sys.path.insert(0, p)
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    # the stubs package `common`/`utils` must win over anything else of that name
    for name in ('common', 'common.interventions', 'utils', 'utils.perf'):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, '__file__', '').startswith(os.path.join(_HERE, 'stubs')):
            del sys.modules[name]
    from cythonsim import main as model
    return model


def make_context(variables=None, area=None, scenario=None, seed=None, age_count_override=None,
                 interventions=None):
    import pandas as pd
    from reina_b200 import inputs
    model = load_model()
    v = variables or inputs.default_variables()
    args = inputs.build_context_args(v, area=area, age_count_override=age_count_override)
    pp = args['population_params']
    pp['age_structure'] = pd.Series(pp['age_structure'])
    pp['contacts_per_day'] = pd.DataFrame(
        pp['contacts_per_day'], columns=['place_type', 'participant_age', 'contact_age', 'contacts'])
    if seed is not None:
        args['random_seed'] = seed
    ctx = model.Context(**args)
    ivs = interventions if interventions is not None else inputs.active_interventions(v, scenario)
    for iv in ivs:
        ctx.add_intervention(iv)
    return ctx


def state_row(s):
    """One generate_state() dict -> flat float64 vector (order: series_names())."""
    row = [float(np.sum(s[a])) for a in POP_SERIES]
    row += [float(s[a]) for a in SCALAR_SERIES]
    row += [float(s['daily_contacts'][p]) for p in PLACES]
    row += [float(x) for x in s['infected_by_variant'].values()]
    return row


def series_names(variant_names=('wild-type', 'b1.1.7')):
    return (POP_SERIES + SCALAR_SERIES + ['exposures_%s' % p for p in PLACES]
            + ['infected_by_variant_%s' % v for v in variant_names])


def run(days=180, keep_age_groups=False, **kw):
    """Returns (series[days, n_series], seconds spent inside iterate(), by_group or None)."""
    ctx = make_context(**kw)
    rows = []
    groups = []
    t_iter = 0.0
    for _ in range(days):
        s = ctx.generate_state()
        rows.append(state_row(s))
        if keep_age_groups:
            groups.append(np.stack([np.asarray(s[a]) for a in POP_SERIES]))
        t0 = time.perf_counter()
        ctx.iterate()
        t_iter += time.perf_counter() - t0
    return np.asarray(rows), t_iter, (np.asarray(groups) if keep_age_groups else None)


def _worker(job):
    kw, days = job
    rows, t, _ = run(days=days, **kw)
    return rows, t


def run_ensemble(seeds, days=180, processes=None, **kw):
    """Monte-Carlo ensemble over seeds with a process pool (mirrors run_monte_carlo,
    calc/simulation.py:376-377).  Returns (series[n_seeds, days, n_series], iterate seconds per seed,
    wall seconds)."""
    import multiprocessing as mp
    jobs = [(dict(kw, seed=int(s)), days) for s in seeds]
    processes = processes or min(len(jobs), os.cpu_count() or 1)
    t0 = time.perf_counter()
    if processes == 1:
        res = [_worker(j) for j in jobs]
    else:
        with mp.get_context('fork').Pool(processes) as pool:
            res = pool.map(_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return np.stack([r[0] for r in res]), np.array([r[1] for r in res]), wall


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--area', default='HUS')
    ap.add_argument('--days', type=int, default=180)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--scenario', default=None)
    a = ap.parse_args()
    rows, t, _ = run(days=a.days, area=a.area, seed=a.seed, scenario=a.scenario)
    names = series_names()
    last = dict(zip(names, rows[-1]))
    print('iterate() total %.2f s' % t)
    for k in ('all_infected', 'infected', 'dead', 'in_ward', 'in_icu', 'all_detected', 'recovered'):
        print('%-14s %d' % (k, last[k]))
    print('peak infected %d on day %d' % (rows[:, 2].max(), rows[:, 2].argmax()))
