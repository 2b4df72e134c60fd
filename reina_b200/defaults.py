"""Default model variables and scenario intervention lists (configuration DATA, not code).

Values restate the reference's configuration so the benchmark configs of BASELINE.json can be run
without the reference tree:  variables.py:227-435 (VARIABLE_DEFAULTS), scenarios.py:89-200
(scenario intervention lists).  When this package is plugged into the real `calc.simulation`, the
reference's own `variables` module is used instead and this file is not read.
"""
import copy

_DECADES = list(range(0, 100, 10))


def _by_decade(vals):
    return [[a, v] for a, v in zip(_DECADES, vals)]


VARIABLE_DEFAULTS = {
    'area_name': 'HUS',
    'country': 'FI',
    'max_age': 100,
    'simulation_days': 565,
    'start_date': '2020-02-18',
    'hospital_beds': 2600,
    'icu_units': 300,
    'p_mask_protects_wearer': 10.0,
    'p_mask_protects_others': 70.0,
    'infectiousness_multiplier': 0.55,
    'p_susceptibility': _by_decade([34.0, 67.0, 100.0, 100.0, 100.0, 100.0, 124.0, 147.0, 147.0, 147.0]),
    'p_asymptomatic_infection': 0.8,
    'p_symptomatic': _by_decade([50.0, 55.0, 60.0, 65.0, 70.0, 75.0, 80.0, 85.0, 90.0, 90.0]),
    'p_severe': _by_decade([0.05, 0.165, 0.72, 2.08, 3.43, 7.65, 13.28, 20.655, 24.57, 24.57]),
    'p_critical': _by_decade([0.003, 0.008, 0.036, 0.104, 0.216, 0.933, 3.639, 8.923, 17.42, 17.42]),
    'p_fatal': _by_decade([0.002, 0.002, 0.01, 0.032, 0.098, 0.265, 0.766, 2.439, 8.292, 16.19]),
    'p_death_outside_hospital': _by_decade([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 6.0, 50.0, 55.0]),
    'p_hospital_death_no_beds': 20.0,
    'p_icu_death_no_beds': 100.0,
    'mean_incubation_duration': 5.1,
    'mean_duration_from_onset_to_death': 18.8,
    'mean_duration_from_onset_to_recovery': 21.0,
    'ratio_of_duration_before_hospitalisation': 30.0,
    'ratio_of_duration_in_ward': 15.0,
    'imported_infection_ages': [[0, 15.0], [20, 40.0], [40, 40.0], [60, 5.0], [70, 0]],
    # variables.py:362-364: the unmeasurable part of the initial population condition (calc/datasets.py:166-168)
    'incubating_at_simulation_start': 0,
    'ill_at_simulation_start': 0,
    'recovered_at_simulation_start': 0,
    'interventions': [
        ['test-all-with-symptoms', '2020-02-20'],
        ['test-only-severe-symptoms', '2020-03-15', 25],
        ['test-only-severe-symptoms', '2020-03-30', 50],
        ['test-only-severe-symptoms', '2020-04-15', 70],
        ['test-with-contact-tracing', '2020-06-15', 30],
        ['test-with-contact-tracing', '2020-09-15', 30],
        ['limit-mobility', '2020-03-15', 80, 0, 70, 'other'],
        ['limit-mobility', '2020-08-15', 50, 0, 70, 'other'],
        ['limit-mobility', '2020-04-01', 5],
        ['limit-mobility', '2020-05-01', 20],
        ['limit-mobility', '2020-07-01', 10],
        ['limit-mobility', '2020-09-01', 10],
        ['limit-mobility', '2020-09-15', 10],
        ['limit-mobility', '2020-10-01', 0],
        ['wear-masks', '2020-07-01', 80, 65, None, None],
        ['limit-mobility', '2020-03-12', 0, 7, 12, 'school'],
        ['limit-mobility', '2020-04-01', 100, 19, None, 'school'],
        ['limit-mobility', '2020-05-30', 100, 7, 12, 'school'],
        ['limit-mobility', '2020-05-30', 100, 13, 15, 'school'],
        ['limit-mobility', '2020-05-30', 100, 16, 18, 'school'],
        ['limit-mobility', '2020-08-12', 0, 7, 12, 'school'],
        ['limit-mobility', '2020-08-12', 0, 13, 15, 'school'],
        ['limit-mobility', '2020-08-12', 0, 16, 18, 'school'],
        ['limit-mobility', '2020-08-12', 20, 19, None, 'school'],
        ['import-infections', '2020-02-22', 20],
        ['import-infections', '2020-03-05', 50],
        ['import-infections', '2020-03-07', 80],
        ['import-infections', '2020-03-09', 120],
        ['import-infections', '2020-03-11', 80],
        ['import-infections', '2020-03-13', 20],
        ['import-infections', '2020-03-15', 20],
        ['import-infections-weekly', '2020-07-01', 50],
        ['import-infections', '2020-08-15', 50],
        ['import-infections', '2020-09-01', 100],
        ['import-infections', '2020-09-07', 100],
        ['import-infections', '2020-09-15', 100],
        ['import-infections', '2020-10-01', 50],
        ['import-infections', '2020-10-15', 100],
        ['import-infections', '2020-11-01', 100],
        ['import-infections', '2020-11-15', 100],
    ],
    # variables.py:413-417,434-435: the variant is 65 % more infectious than wild-type
    'variants': [{'name': 'b1.1.7', 'infectiousness_multiplier': 0.55 * 1.65}],
    'scenarios': [{'id': 'default'}],
    'active_scenario': 'default',
    'random_seed': 0,
}

# scenarios.py:89-161 -- extra intervention tuples appended to the default list (config #3)
SCENARIO_INTERVENTIONS = {
    'default': [],
    'summer-boogie': [['limit-mobility', '2020-05-15', 30]],
    'mitigation': (
        [[kind, date, n]
         for date in ('2020-06-30', '2020-07-15', '2020-07-30', '2020-08-15', '2020-08-30')
         for kind, n in (('build-new-icu-units', 150), ('build-new-hospital-beds', 300))]
        + [['limit-mobility', d, v] for d, v in (
            ('2020-06-01', 30), ('2020-07-01', 40), ('2020-08-01', 30), ('2020-09-15', 40),
            ('2020-10-15', 30), ('2020-12-15', 20), ('2021-01-15', 5), ('2021-02-15', 0))]
    ),
    'hammer-and-dance': (
        [['test-with-contact-tracing', d, v] for d, v in (
            ('2020-05-01', 30), ('2020-06-01', 40), ('2020-07-01', 50), ('2020-08-01', 60))]
        + [['limit-mobility', d, v] for d, v in (
            ('2020-05-01', 30), ('2020-06-24', 25), ('2020-08-15', 10), ('2020-12-06', 15))]
    ),
}


def default_variables(**overrides):
    v = copy.deepcopy(VARIABLE_DEFAULTS)
    v.update(overrides)
    return v


def scenario_interventions(scenario='default', base=None):
    """Intervention tuple list for a scenario (list concatenation; the reference's own
    Scenario.apply is broken).  'looser-restrictions-to-start-with' halves
    every limit-mobility value of the base list (scenarios.py:181-191)."""
    base = copy.deepcopy(VARIABLE_DEFAULTS['interventions'] if base is None else base)
    if scenario == 'looser-restrictions-to-start-with':
        out = []
        for iv in base:
            iv = list(iv)
            if iv[0] == 'limit-mobility':
                iv[2] = iv[2] // 2
            out.append(iv)
        return out
    return base + copy.deepcopy(SCENARIO_INTERVENTIONS[scenario])
