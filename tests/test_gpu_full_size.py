"""Full-size checks on the GPU (BASELINE configs 1-4 sizes): statistical parity of the CUDA engine with the
UNMODIFIED reference engine (golden ensemble statistics, >= 64 seeds per side, every daily series within 3 standard
errors -- north_star's acceptance rule), and size-independent properties the domain offers."""
import os

import numpy as np
import pytest

import helpers
from test_oracle_pinned import zscores

pytestmark = pytest.mark.gpu
GOLD = os.path.join(helpers.ROOT, 'tests', 'golden')


@pytest.mark.parametrize('gold_name,area,scenario', [
    ('hus_default', 'HUS', None),                                  # BASELINE configs[1]
    ('varsinais_suomi_default', 'Varsinais-Suomi', None),          # configs[0]
    ('hus_hammer_and_dance', 'HUS', 'hammer-and-dance'),           # configs[2]: contact tracing 30 -> 60 %
    ('hus_mitigation', 'HUS', 'mitigation'),                       # configs[2]: capacity building + mobility
    ('hus_summer_boogie', 'HUS', 'summer-boogie'),
    ('hus_looser_restrictions', 'HUS', 'looser-restrictions-to-start-with'),   # configs[2]: every limit-mobility value halved
    ('hus_initial_state', 'HUS', None),                            # Population.set_initial_state, start 2020-04-01
])
def test_ensemble_statistics_match_reference(cuda_lib, gold_name, area, scenario):
    gold = np.load(os.path.join(GOLD, 'ref_ensemble_%s.npz' % gold_name))
    variables = None
    if gold_name == 'hus_initial_state':
        from test_oracle_pinned import INITIAL_STATE_VARIABLES
        variables = helpers.inputs.default_variables()
        variables.update(INITIAL_STATE_VARIABLES)
    ctx = helpers.make_context(cuda_lib, area=area, scenario=scenario, seed=424242, n_replicas=64, max_days=181, variables=variables)
    ctx.run(180)
    mine = helpers.series_matrix(ctx)
    ctx.close()
    names = list(gold['names'])
    assert names == helpers.series_names()
    z, exact = zscores(mine, gold)
    assert not exact.any(), 'deterministic series differ: %s' % sorted({names[j] for j in np.argwhere(exact)[:, 1]})
    frac = (np.abs(z) > 3).mean()
    worst = np.unravel_index(np.abs(z).argmax(), z.shape)
    print('%s: fraction of (day, series) cells beyond 3 SE %.4f, worst |z| %.2f (%s day %d)'
          % (gold_name, frac, np.abs(z).max(), names[worst[1]], worst[0]))
    assert frac < 0.015
    assert np.abs(z).max() < 5.0
    for s in ('all_infected', 'dead', 'all_detected', 'recovered', 'cum_icu', 'in_ward'):
        assert abs(z[-1, names.index(s)]) < 3.0, (s, z[-1, names.index(s)])


def test_full_size_properties(cuda_lib):
    """HUS, 1,685,983 agents, 180 days: conservation laws, monotone cumulative counters, capacity accounting,
    determinism, and replica r of an ensemble == the single run with seed + r."""
    N = 1685983
    ctx = helpers.make_context(cuda_lib, seed=5, n_replicas=3, max_days=181)
    ctx.run(180)
    rows = ctx.series(0, 180)
    m = helpers.series_matrix(ctx)
    names = helpers.series_names()
    col = {n: m[:, :, i] for i, n in enumerate(names)}
    assert np.all(col['susceptible'] + col['infected'] + col['recovered'] + col['dead'] == N)
    assert np.all(col['all_infected'] == N - col['susceptible'])
    for cum in ('all_infected', 'all_detected', 'dead', 'recovered', 'cum_icu', 'non_hospital_deaths'):
        assert np.all(np.diff(col[cum], axis=1) >= 0), cum
    assert np.all(col['in_ward'] + col['available_hospital_beds'] == 2600)
    assert np.all(col['in_icu'] + col['available_icu_units'] == 300)
    assert np.all(col['available_hospital_beds'] >= 0) and col['available_hospital_beds'].min() == 0   # saturates
    assert np.all(col['non_hospital_deaths'] <= col['dead'])
    assert np.all(col['new_infections'][:, 1:] <= np.diff(col['all_infected'], axis=1) + 1e-9)
    expo = sum(col['exposures_%s' % p] for p in helpers.PLACES)
    assert np.all(expo == col['exposed_per_day'])
    agents = ctx._engine.read_agents(1)
    states = np.bincount(agents['state'], minlength=7)
    assert states.sum() == N
    ctx2 = helpers.make_context(cuda_lib, seed=6, max_days=181)
    ctx2.run(180)
    assert np.array_equal(ctx2.series(0, 180)[0], rows[1])              # replica 1 == single run with seed 5 + 1
    ctx2.reset(6)
    ctx2.run(180)
    assert np.array_equal(ctx2.series(0, 180)[0], rows[1])              # bit-reproducible
    a2 = ctx2._engine.read_agents(0)
    assert np.array_equal(a2, agents)
    # end state of the agents agrees with the last stats row's successor (one more snapshot)
    s = ctx2.generate_state()
    assert s['dead'].sum() == (a2['state'] == 6).sum() and s['recovered'].sum() == (a2['state'] == 5).sum()
    assert s['infected'].sum() == ((a2['state'] >= 1) & (a2['state'] <= 4)).sum()


def test_simulate_individuals_frames(cuda_lib):
    """The calc.simulation-shaped entry point over the GPU engine (configs[0]: Varsinais-Suomi)."""
    from reina_b200 import inputs, simulation
    v = inputs.default_variables(simulation_days=120, area_name='Varsinais-Suomi', random_seed=0)
    df, adf = simulation.simulate_individuals(v)
    assert list(df.columns) == simulation.POP_ATTRS + simulation.STATE_ATTRS + simulation.EXPOSURES_ATTRS + ['us_per_infected']
    assert df.shape[0] == 120 and adf.shape == (120, 12 * 9)
    assert df['susceptible'].iloc[0] == 479861 and df['all_infected'].iloc[-1] > 1000
    assert str(df.index[0].date()) == '2020-02-18'
