"""Caller side of the hot path: the entry points of `calc.simulation` over the CUDA engine.

Mirrors calc/simulation.py (reference): `simulate_individuals` (:148-290) returns the same two DataFrames --
`df` (one row per day: POP_ATTRS + STATE_ATTRS + EXPOSURES_ATTRS + 'us_per_infected', DatetimeIndex) and `adf`
(per age group, MultiIndex columns (attr, age_group)) -- `sample_model_parameters` (:301-346) and a working
`run_monte_carlo` (the reference's, :349-385, treats the (df, adf) tuple as a DataFrame and its scenario loader is
broken).  The per-day Python/pandas loop of the reference (1.7 ms per day, SURVEY
section 6) would dwarf a GPU day, so results are fetched once per run (or once per `callback_day_interval` days when
a step_callback is given) and the DataFrames are built in one go.

    AREA_NAME=Varsinais-Suomi python -m reina_b200.simulation [--days 180] [--seed 0] [--scenario mitigation]
prints the reference CLI's table (calc/simulation.py:408-446).  AREA_NAME is the README's knob
(README.md:78-84); the reference code itself never reads it, here it sets variables['area_name'].
"""
import os
import time
from datetime import date

import numpy as np

from . import _abi, inputs, model

# calc/simulation.py:17-47
POP_ATTRS = ['susceptible', 'vaccinated', 'infected', 'detected', 'all_detected', 'in_ward', 'in_icu', 'dead',
             'non_hospital_deaths', 'recovered', 'all_infected', 'new_infections']
EXPOSURES_ATTRS = ['exposures_home', 'exposures_work', 'exposures_school', 'exposures_transport',
                   'exposures_leisure', 'exposures_other']
STATE_ATTRS = ['exposed_per_day', 'available_hospital_beds', 'available_icu_units', 'total_icu_units',
               'ct_cases_per_day', 'r', 'mobility_limitation']


class ExecutionInterrupted(Exception):   # calc/__init__.py
    pass


def make_context(variables=None, n_replicas=1, device=0, scenario=None, **kw):
    """calc/simulation.py:151-180."""
    v = variables or inputs.default_variables()
    args = inputs.build_context_args(v)
    ctx = model.Context(n_replicas=n_replicas, device=device, max_days=max(v['simulation_days'] + 1, 2), **args, **kw)
    for iv in inputs.active_interventions(v, scenario):
        ctx.add_intervention(iv)
    return ctx


def rows_to_frames(ctx, rows, start_date, us_per_infected=None):
    """Stats rows of ONE replica [days, row_len] -> (df, adf) exactly as calc/simulation.py:183-290 shapes them."""
    import pandas as pd
    days = rows.shape[0]
    G = len(ctx.age_group_labels)
    nA = len(_abi.ATTRS)
    sc = rows[:, nA * G:]
    S = {name: sc[:, i] for i, name in enumerate(_abi.SCALARS)}
    date_index = pd.date_range(date.fromisoformat(start_date), periods=days)
    cols = {}
    ag = np.empty((days, len(POP_ATTRS), G), dtype='i')
    for k, attr in enumerate(POP_ATTRS):
        i = _abi.ATTRS.index(attr)
        ag[:, k, :] = rows[:, i * G:(i + 1) * G]
        cols[attr] = ag[:, k, :].sum(axis=1)
    inf, tor = S['total_infections'].astype(np.float64), S['total_infectors'].astype(np.float64)
    cols['exposed_per_day'] = S['exposed_per_day']
    cols['available_hospital_beds'] = S['available_hospital_beds']
    cols['available_icu_units'] = S['available_icu_units']
    cols['total_icu_units'] = S['total_icu_units']
    cols['ct_cases_per_day'] = S['ct_cases_per_day']
    cols['r'] = np.where(tor > 5, inf / np.maximum(tor, 1), 0.0)             # main.pyx:1817
    cols['mobility_limitation'] = [ctx._epoch_mobility.get(int(e), 0.0) for e in S['table_epoch']]
    place_idx = {'home': 0, 'work': 1, 'school': 2, 'transport': 3, 'leisure': 4, 'other': 5}
    for name in EXPOSURES_ATTRS:
        cols[name] = sc[:, _abi.RB_S_CONTACTS0 + place_idx[name.split('_', 1)[1]]]
    cols['us_per_infected'] = np.zeros(days) if us_per_infected is None else us_per_infected
    df = pd.DataFrame(cols, index=date_index, columns=POP_ATTRS + STATE_ATTRS + EXPOSURES_ATTRS + ['us_per_infected'])
    adf = pd.DataFrame(
        ag.flatten(),
        index=pd.MultiIndex.from_product([date_index, POP_ATTRS, ctx.age_group_labels], names=['date', 'attr', 'age_group']),
        columns=['pop'])
    adf = adf.unstack('attr').unstack('age_group')
    adf.columns = adf.columns.droplevel()
    return df, adf


def simulate_individuals(variables=None, step_callback=None, callback_day_interval=1, replica=0, context=None, **_ignored):
    """calc/simulation.py:148-290.  Row d is the state BEFORE the d-th iterate(), as in the reference."""
    v = variables or inputs.default_variables()
    ctx = context or make_context(v)
    days = v['simulation_days']
    t0 = time.perf_counter()
    if step_callback is None:
        ctx.run(days)
    else:
        done = 0
        while done < days:
            n = min(max(1, callback_day_interval), days - done)
            ctx.run(n)
            done += n
            part, _ = rows_to_frames(ctx, ctx.series(0, done)[replica], v['start_date'])
            part = part.reindex(part.index.union(
                __import__('pandas').date_range(part.index[0], periods=days)))   # future days are NaN rows
            if not step_callback(part):
                raise ExecutionInterrupted()
    rows = ctx.series(0, days)[replica]
    G = len(ctx.age_group_labels)
    infected = rows[:, _abi.ATTRS.index('infected') * G:(_abi.ATTRS.index('infected') + 1) * G].sum(axis=1)
    ms_per_day = (time.perf_counter() - t0) * 1e3 / max(days, 1)
    us = np.where(infected > 0, ms_per_day * 1000 / np.maximum(infected, 1), 0.0)
    return rows_to_frames(ctx, rows, v['start_date'], us)


def sample_model_parameters(what, age, severity=None, variables=None):
    """calc/simulation.py:301-346 (returns the normalised value counts; the reference's plotting is left out)."""
    import pandas as pd
    v = variables or inputs.default_variables()
    args = inputs.build_context_args(v, age_count_override=np.ones(v['max_age'] + 1, dtype=np.int64))
    args['healthcare_params'] = dict(hospital_beds=0, icu_units=0)
    args['start_date'] = '2020-01-01'
    args.pop('random_seed')
    ctx = model.Context(**args)
    samples = ctx.sample(what, age, severity)
    if what == 'infectiousness':
        s = pd.Series(index=samples['day'], data=samples['val'])
        return s[s != 0].sort_index()
    c = pd.Series(samples).value_counts().sort_index()
    if what == 'symptom_severity':
        c.index = c.index.map(model.SEVERITY_TO_STR)
    return c / c.sum()


def run_monte_carlo(scenario_name='default', n_seeds=1000, seed0=0, variables=None, replicas_per_launch=64,
                    device=0, csv_path=None, bands=None, context=None):
    """Working replacement of calc/simulation.py:365-385: `n_seeds` runs of one scenario, `replicas_per_launch`
    at a time on the GPU; returns the long DataFrame (date, columns..., run, scenario) and writes
    reina_<scenario>.csv like the reference does.  With `bands=(5, 50, 95)` also returns the percentile bands of every
    column per day: {q: DataFrame[date x columns]}."""
    import pandas as pd
    v = variables or inputs.default_variables()
    scenario = None if scenario_name in (None, 'default') else scenario_name
    ctx = context or make_context(v, n_replicas=replicas_per_launch, device=device, scenario=scenario)
    replicas_per_launch = ctx.n_replicas
    days = v['simulation_days']
    dfs = []
    for first in range(seed0, seed0 + n_seeds, replicas_per_launch):
        ctx.reset(first)
        ctx.run(days)
        rows = ctx.series(0, days)
        for r in range(min(replicas_per_launch, seed0 + n_seeds - first)):
            df, _ = rows_to_frames(ctx, rows[r], v['start_date'])
            df['run'] = first + r
            dfs.append(df)
    out = None
    if bands:
        from . import ensemble
        cols = [c for c in dfs[0].columns if c not in ('run', 'us_per_infected')]
        cube = np.stack([d[cols].to_numpy(dtype=np.float64) for d in dfs])          # [run, day, column]
        out = {q: pd.DataFrame(p, index=dfs[0].index, columns=cols) for q, p in ensemble.percentile_bands(cube, bands).items()}
    df = pd.concat(dfs)
    df.index.name = 'date'
    df = df.reset_index()
    df['scenario'] = scenario_name
    if csv_path is not False:
        df.to_csv(csv_path or 'reina_%s.csv' % scenario_name, index=False)
    return (df, out) if bands else df


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--days', type=int, default=None)
    ap.add_argument('--seed', type=int, default=None)
    ap.add_argument('--scenario', default=None)
    a = ap.parse_args(argv)
    v = inputs.default_variables()
    if os.environ.get('AREA_NAME'):
        v['area_name'] = os.environ['AREA_NAME']
    if a.days:
        v['simulation_days'] = a.days
    if a.seed is not None:
        v['random_seed'] = a.seed
    ctx = make_context(v, scenario=a.scenario)
    cols = POP_ATTRS + ['ct_cases_per_day', 'r', 'exposures', 'us_per_infected']
    print('%-10s' % 'day' + ''.join('%15s' % c for c in cols))
    df, adf = simulate_individuals(v, context=ctx)
    for day, rec in df.iterrows():
        s = '%-12s' % day.date().isoformat() + ''.join('%15d' % rec[c] for c in POP_ATTRS)
        s += '%15d' % rec['ct_cases_per_day'] + '%13.2f' % rec['r']
        s += '%15d' % sum(rec[c] for c in EXPOSURES_ATTRS)
        s += '%13.2f' % rec['us_per_infected'] if rec['infected'] else ''
        print(s)
    print(adf)


if __name__ == '__main__':
    main()
