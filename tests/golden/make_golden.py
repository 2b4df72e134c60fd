"""Generate golden ensemble statistics from the UNMODIFIED reference engine (oracle/_ref).

Run in the build container (where oracle/build_ref.sh can build the reference):
    python tests/golden/make_golden.py [--seeds 64] [--configs ...]
Writes tests/golden/ref_ensemble_<config>.npz with, per (day, series): mean, sample std (ddof=1)
and n over the seeds, plus the series names.  These files pin the statistical parity tests
(north_star: every daily series within 3 standard errors over >= 64 seeds per side).

The reference has no golden vectors of its own for this path ("parity
unpinned"), so outputs of the reference itself are the pin.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402

CONFIGS = {
    # name: (area, scenario, days)
    'varsinais_suomi_default': ('Varsinais-Suomi', None, 180),
    'hus_default': ('HUS', None, 180),
    'hus_hammer_and_dance': ('HUS', 'hammer-and-dance', 180),
    'hus_mitigation': ('HUS', 'mitigation', 180),
    'hus_summer_boogie': ('HUS', 'summer-boogie', 180),
    'hus_looser_restrictions': ('HUS', 'looser-restrictions-to-start-with', 180),
    # Population.set_initial_state (main.pyx:1452-1516): a start date the HUS case file holds (9 dead, 32 in ICU, 52 in
    # ward, 1200 confirmed) + the example values of variables.py:212-214 for the unmeasurable part
    'hus_initial_state': ('HUS', None, 180),
    # every intervention type (imports of both variants, weekly trickle, all testing modes, contact tracing at three
    # efficiencies, masks, age / place mobility limits, three vaccination updates, capacity building) on an 80,000-agent
    # HUS-shaped population with a tiny hospital that saturates: tests/helpers.py stress_interventions(), 256 seeds
    'hus80k_every_intervention': ('HUS', None, 120),
}
VARIABLE_OVERRIDES = {
    'hus_initial_state': dict(start_date='2020-04-01', incubating_at_simulation_start=150, ill_at_simulation_start=50,
                              recovered_at_simulation_start=1000),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', type=int, default=64)
    ap.add_argument('--seed0', type=int, default=1000)
    ap.add_argument('--processes', type=int, default=None)
    ap.add_argument('--configs', nargs='*', default=list(CONFIGS))
    ap.add_argument('--suffix', default='', help="appended to the file name, e.g. _n256 for a higher-power ensemble")
    a = ap.parse_args()
    for name in a.configs:
        area, scenario, days = CONFIGS[name]
        t0 = time.time()
        seeds = np.arange(a.seed0, a.seed0 + a.seeds)
        variables = None
        if name in VARIABLE_OVERRIDES:
            from reina_b200 import inputs
            variables = inputs.default_variables()
            variables.update(VARIABLE_OVERRIDES[name])
        extra = {}
        if name == 'hus80k_every_intervention':
            sys.path.insert(0, os.path.join(ROOT, 'tests'))
            import helpers
            from reina_b200 import inputs
            seeds = np.arange(a.seed0, a.seed0 + max(a.seeds, 256))
            variables = inputs.default_variables()
            variables['hospital_beds'], variables['icu_units'] = 25, 3
            extra = dict(age_count_override=helpers.small_population(80000), interventions=helpers.stress_interventions())
        series, t_iter, wall = ref_harness.run_ensemble(seeds, days=days, processes=a.processes,
                                                        area=area, scenario=scenario, variables=variables, **extra)
        out = os.path.join(HERE, 'ref_ensemble_%s%s.npz' % (name, a.suffix))
        np.savez_compressed(
            out, mean=series.mean(axis=0), std=series.std(axis=0, ddof=1), n=len(seeds),
            names=np.array(ref_harness.series_names()), seeds=seeds,
            iterate_seconds=t_iter, area=area, scenario=str(scenario), days=days)
        print('%s: %d seeds in %.0f s (mean iterate %.1f s/seed) -> %s' % (
            name, len(seeds), time.time() - t0, t_iter.mean(), out), flush=True)


if __name__ == '__main__':
    main()
