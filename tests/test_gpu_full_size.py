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


HIGH_POWER = [
    # golden file, area, scenario, base seed of the 256 replicas, variable overrides
    ('hus_default_n256', 'HUS', None, 909090, None),                       # BASELINE configs[1] / [3]
    ('hus_hammer_and_dance_n256', 'HUS', 'hammer-and-dance', 919191, None),   # configs[2]: contact tracing 30 -> 60 % + mobility
    ('hus_mitigation_n256', 'HUS', 'mitigation', 929292, None),            # configs[2]: capacity building + mobility limits
    ('hus_summer_boogie_n256', 'HUS', 'summer-boogie', 939393, None),
    ('hus_looser_restrictions_n256', 'HUS', 'looser-restrictions-to-start-with', 949494, None),   # configs[2]: every limit-mobility value halved
    ('hus_initial_state_n256', 'HUS', None, 959595, 'initial_state'),      # Population.set_initial_state, start 2020-04-01
    ('varsinais_suomi_default_n256', 'Varsinais-Suomi', None, 969696, None),   # configs[0]
]


def high_power_variables(key):
    if key is None:
        return None
    from test_oracle_pinned import INITIAL_STATE_VARIABLES
    v = helpers.inputs.default_variables()
    v.update(INITIAL_STATE_VARIABLES)
    return v


def high_power_report(mine, gold_name):
    """256 runs of ours (rows [256, days, series]) against 256 seeds of the UNMODIFIED reference engine: the statistics the
    test asserts on, shared with tests/golden/precheck_high_power.py (which computes `mine` with the sequential oracle
    from the very seeds the GPU test uses -- bit-identical rows -- so the thresholds are known to hold before a GPU run)."""
    gold = np.load(os.path.join(GOLD, 'ref_ensemble_%s.npz' % gold_name))
    assert int(gold['n']) == 256 and mine.shape[0] == 256
    names = list(gold['names'])
    assert names == helpers.series_names()
    z, exact = zscores(mine, gold)
    assert not exact.any(), 'deterministic series differ: %s' % sorted({names[j] for j in np.argwhere(exact)[:, 1]})
    frac3, frac2 = (np.abs(z) > 3).mean(), (np.abs(z) > 2).mean()
    worst = np.unravel_index(np.abs(z).argmax(), z.shape)
    print('%s, 256 + 256 seeds: cells beyond 3 SE %.4f, beyond 2 SE %.4f (a normal gives 0.0027 / 0.0455), worst |z| %.2f (%s day %d)'
          % (gold_name, frac3, frac2, np.abs(z).max(), names[worst[1]], worst[0]))
    assert frac3 < 0.01 and frac2 < 0.10
    assert np.abs(z).max() < 4.5
    m, g = mine.mean(0)[-1], gold['mean'][-1]
    for s in ('all_infected', 'dead', 'all_detected', 'recovered', 'cum_icu'):
        j = names.index(s)
        rel = (m[j] - g[j]) / g[j]
        print('   day 180 %-14s ours %12.1f  reference %12.1f  (%+.2f %%, z %+.2f)' % (s, m[j], g[j], 100 * rel, z[-1, j]))
        assert abs(z[-1, j]) < 3.0 and abs(rel) < 0.025, (s, rel, z[-1, j])


@pytest.mark.parametrize('gold_name,area,scenario,seed,variables', HIGH_POWER)
def test_high_power_statistics(cuda_lib, gold_name, area, scenario, seed, variables):
    """Higher statistical power than the 64-seed tests above: 256 CUDA replicas (the bench's own ensemble size) against 256
    seeds of the UNMODIFIED reference engine (tests/golden/make_golden.py --seeds 256 --seed0 50000 --suffix _n256).  With
    64 + 64 seeds a modelling error of ~1.5 % in a total hides inside 3 SE; at 256 + 256 one standard error of the day-180
    totals is 0.5-0.75 %, which is what keeps the accelerations shared by oracle and engine (thinning, tabulated contact
    count, 24-bit row uniforms) honest -- for every configuration the 64-seed tests above cover."""
    ctx = helpers.make_context(cuda_lib, area=area, scenario=scenario, seed=seed, n_replicas=256, max_days=181,
                               variables=high_power_variables(variables))
    ctx.run(180)
    mine = helpers.series_matrix(ctx)
    ctx.close()
    high_power_report(mine, gold_name)


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


# ---------------------------------------------------------------------------------------------------
# bit-exact at the FULL size and in the launch geometry bench.py times
# ---------------------------------------------------------------------------------------------------
N_HUS = 1685983


def test_full_size_single_seed_bit_exact(cuda_lib, oracle_lib):
    """BASELINE configs[1] literally: HUS, 1,685,983 agents x 180 days, one seed -- CUDA == sequential oracle in every
    stats row, every agent field, the test queue and the capacity counters (full-size bucket counts, list capacities
    and age-block walks; the single-seed launch geometry)."""
    kw = dict(seed=31337, max_days=181)
    gpu = helpers.make_context(cuda_lib, **kw)
    cpu = helpers.make_context(oracle_lib, **kw)
    assert gpu.n_agents == N_HUS
    gpu.run(180)
    cpu.run(180)
    rep = helpers.diff_report(gpu, cpu, 180)
    assert not rep, '\n'.join(rep)
    gpu.close()


def test_bench_geometry_bit_exact(cuda_lib, oracle_lib):
    """The exact configuration `python bench.py` times -- 256 HUS replicas, the engine's DEFAULT replica groups / grids /
    streams for that count -- is bit-identical, for replicas 0, 100 and 255, to three single-seed CUDA runs AND to the
    sequential oracle (stats rows + every agent field)."""
    R, seed0 = 256, 900000
    ens = helpers.make_context(cuda_lib, seed=seed0, n_replicas=R, max_days=181)
    assert ens.n_agents == N_HUS
    for k in ('RB_GROUPS', 'RB_GROUP_WAVE_PCT', 'RB_WIDE_CTAS', 'RB_WIDE_MIN'):
        assert k not in os.environ, 'this test must run with the default launch geometry'
    ens.run(180)
    rows = ens.series(0, 180)
    one = helpers.make_context(cuda_lib, seed=seed0, max_days=181)
    for r in (0, 100, 255):
        cpu = helpers.make_context(oracle_lib, seed=seed0 + r, max_days=181)
        cpu.run(180)
        ref = cpu.series(0, 180)[0]
        bad = np.argwhere(rows[r] != ref)
        assert len(bad) == 0, 'replica %d of the ensemble differs from the oracle first at (day, col) %s' % (r, bad[0])
        ca = cpu._engine.read_agents(0)
        assert np.array_equal(ens._engine.read_agents(r), ca), 'replica %d: agent fields differ from the oracle' % r
        assert np.array_equal(np.sort(ens._engine.read_queue(r)), np.sort(cpu._engine.read_queue(0)))
        assert np.array_equal(ens._engine.read_available(r), cpu._engine.read_available(0))
        one.reset(seed0 + r)
        one.run(180)
        assert np.array_equal(one.series(0, 180)[0], rows[r]), 'replica %d differs from the single-seed CUDA run' % r
        assert np.array_equal(one._engine.read_agents(0), ca)
    ens.close()
    one.close()


def test_monte_carlo_driver_on_cuda(cuda_lib, oracle_lib, tmp_path):
    """simulation.run_monte_carlo on the GPU: 128 seeds at 64 per launch; two sampled runs equal the
    oracle's single-seed runs row for row; percentile bands bracket the median."""
    from reina_b200 import inputs, simulation
    v = inputs.default_variables(simulation_days=90, area_name='Varsinais-Suomi')
    df, bands = simulation.run_monte_carlo('default', n_seeds=128, seed0=5000, variables=v, replicas_per_launch=64,
                                           csv_path=str(tmp_path / 'mc.csv'), bands=(5, 50, 95))
    assert sorted(df['run'].unique()) == list(range(5000, 5128)) and len(df) == 128 * 90
    assert os.path.getsize(tmp_path / 'mc.csv') > 0
    assert (bands[5] <= bands[50]).all().all() and (bands[50] <= bands[95]).all().all()
    for run in (5003, 5100):                          # one from each launch
        cpu = simulation.make_context(v, _library=oracle_lib)
        cpu.reset(run)
        d2, _ = simulation.simulate_individuals(v, context=cpu)
        mine = df[df['run'] == run]
        for col in ('infected', 'all_infected', 'dead', 'exposed_per_day', 'exposures_work', 'available_hospital_beds', 'r'):
            assert np.array_equal(mine[col].to_numpy(dtype=float), d2[col].to_numpy(dtype=float)), (run, col)


def test_simulation_worker_on_cuda(cuda_lib, oracle_lib):
    """serving.SimulationWorker over real CUDA contexts: two seeds through one resident context
    (allocated once, reset in place), results equal the oracle's, and a cancelled job stops."""
    from reina_b200 import inputs, serving, simulation
    w = serving.SimulationWorker(device=0, callback_day_interval=20)
    try:
        jobs = []
        for seed in (11, 12):
            v = inputs.default_variables(simulation_days=60, area_name='Varsinais-Suomi', random_seed=seed)
            jobs.append((seed, v, w.submit(v)))
        for seed, v, job in jobs:
            res = w.wait(job, timeout=300)
            assert res['finished'] and res['error'] is None, res['error']
            cpu = simulation.make_context(v, _library=oracle_lib)
            d2, a2 = simulation.simulate_individuals(v, context=cpu)
            cols = [c for c in d2.columns if c != 'us_per_infected']
            assert np.array_equal(res['total'][cols].to_numpy(dtype=float), d2[cols].to_numpy(dtype=float)), seed
            assert np.array_equal(res['age_groups'].to_numpy(), a2.to_numpy())
        assert w.results(jobs[0][2])['reused_context'] is False and w.results(jobs[1][2])['reused_context'] is True
        v = inputs.default_variables(simulation_days=60, area_name='Varsinais-Suomi', random_seed=13)
        job = w.submit(v)
        w.cancel(job)
        res = w.wait(job, timeout=300)
        assert res['finished'] and res['error'] == 'cancelled'
    finally:
        w.close()
