"""Model inputs for the per-day agent loop, built without pandas/flask/xlrd.

Restates what `calc.simulation.simulate_individuals` hands to `model.Context`
(calc/simulation.py:151-180): age structure, long-form contact list, age groups, disease
parameters and interventions.  Arrays come from reina_b200/data/inputs.json (derived from the
reference's data/ by tools/make_inputs.py).
"""
import json
import os

import numpy as np

from .defaults import default_variables, scenario_interventions  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}

# main.pyx:777-785
DISEASE_PARAMS = (
    'p_susceptibility', 'p_symptomatic', 'p_severe', 'p_critical',
    'p_fatal', 'p_hospital_death_no_beds', 'p_icu_death_no_beds',
    'p_death_outside_hospital', 'p_asymptomatic_infection',
    'infectiousness_multiplier', 'mean_incubation_duration',
    'mean_duration_from_onset_to_death', 'mean_duration_from_onset_to_recovery',
    'ratio_of_duration_before_hospitalisation', 'ratio_of_duration_in_ward',
    'p_mask_protects_wearer', 'p_mask_protects_others', 'variants',
)

# Positional parameter ids of each intervention type (common/interventions.py:159-323)
IV_PARAM_IDS = {
    'test-all-with-symptoms': [],
    'test-only-severe-symptoms': ['mild_detection_rate'],
    'test-with-contact-tracing': ['efficiency'],
    'limit-mobility': ['reduction', 'min_age', 'max_age', 'place'],
    'wear-masks': ['share_of_contacts', 'min_age', 'max_age', 'place'],
    'vaccinate': ['weekly_vaccinations', 'min_age', 'max_age'],
    'import-infections': ['amount', 'variant'],
    'import-infections-weekly': ['weekly_amount'],   # + variant_<name> per configured variant
    'build-new-hospital-beds': ['beds'],
    'build-new-icu-units': ['units'],
}


class Intervention:
    """Duck-type of common.interventions.Intervention as the engine sees it
    (`.type`, `.date`, `.get_param_values()`; main.pyx:1880-1960, 2011-2015)."""

    def __init__(self, type, date, values=None):
        self.type = type
        self.date = date
        self.values = dict(values or {})

    def get_param_values(self):
        return dict(self.values)

    def __repr__(self):
        return 'Intervention(%r, %r, %r)' % (self.type, self.date, self.values)


def iv_tuple_to_obj(iv, variant_names=('b1.1.7',)):
    """common/interventions.py:75-103: positional tuple -> object, `None` values dropped."""
    kind, date = iv[0], iv[1]
    if kind not in IV_PARAM_IDS:
        raise Exception('Invalid intervention type: %s' % kind)
    ids = list(IV_PARAM_IDS[kind])
    if kind == 'import-infections-weekly':
        ids += ['variant_%s' % v for v in variant_names]
    values = {}
    for pid, val in zip(ids, list(iv)[2:]):
        if val is None:
            continue
        values[pid] = val
    return Intervention(kind, date, values)


def _inputs():
    if 'inputs' not in _CACHE:
        with open(os.path.join(_HERE, 'data', 'inputs.json')) as f:
            _CACHE['inputs'] = json.load(f)
    return _CACHE['inputs']


def age_counts(area='HUS'):
    """Population by single-year age 0..100 (calc/datasets.py:47-61 summed over sexes)."""
    return np.asarray(_inputs()['areas'][area], dtype=np.int64)


def synthetic_age_counts(total, area='HUS'):
    """scale a district histogram to `total` agents (largest remainder)."""
    base = age_counts(area).astype(np.float64)
    want = base * (total / base.sum())
    out = np.floor(want).astype(np.int64)
    rem = int(total - out.sum())
    order = np.argsort(-(want - out), kind='stable')
    out[order[:rem]] += 1
    return out


def contacts_long(max_age=100):
    """calc/simulation.py:74-100: one record per (place, single participant age, contact band).

    Returns a list of (place_type, participant_age, (band_lo, band_hi), contacts)."""
    c = _inputs()['contacts']
    bands = [tuple(int(y) for y in b.split('-')) for b in c['contact_bands']]
    out = []
    for bi, band in enumerate(bands):          # pd.melt: column-major
        for row in c['rows']:
            lo, hi = (int(y) for y in row['participant_age'].split('-'))
            for p in range(lo, hi + 1):
                out.append((row['place_type'], p, band, row['contacts'][bi]))
    return out


def make_age_groups(max_age=100):
    """calc/simulation.py:103-116,154-161: labels use an en dash; sorted as np.unique does."""
    age_map = []
    for i in range(0, max_age + 1):
        grp = i // 10
        age_map.append('80+' if grp >= 8 else '%d–%d' % (grp * 10, grp * 10 + 9))
    labels = list(np.unique(age_map))
    return dict(labels=labels, age_indices=[labels.index(x) for x in age_map])


def create_disease_params(variables):
    """calc/simulation.py:50-61: every p_*/ratio_* value is divided by 100 (lists element-wise)."""
    kwargs = {}
    for key in DISEASE_PARAMS:
        val = variables[key]
        if key.startswith('p_') or key.startswith('ratio_'):
            if isinstance(val, list):
                val = [(age, sev / 100) for age, sev in val]
            else:
                val = val / 100
        kwargs[key] = val
    return kwargs


class InitialPopulationCondition:
    """calc/datasets.py:107-135."""

    def __init__(self, dead=0, in_icu=0, in_ward=0, confirmed_cases=0, infected_cases=0, incubating=0, ill=0, recovered=0):
        self.dead, self.in_icu, self.in_ward, self.confirmed_cases = int(dead), int(in_icu), int(in_ward), int(confirmed_cases)
        self.infected_cases, self.incubating, self.ill, self.recovered = int(infected_cases), int(incubating), int(ill), int(recovered)

    def has_initial_state(self):
        return bool(self.dead or self.in_icu or self.in_ward or self.confirmed_cases or self.infected_cases
                    or self.incubating or self.ill or self.recovered)

    def were_ill(self):
        return self.dead + self.recovered + self.in_icu + self.in_ward + self.ill

    def were_incubating(self):
        return self.were_ill() + self.incubating

    def recovered_without_illness(self):
        return self.were_incubating() - self.were_ill()


def initial_population_condition(variables, area=None):
    """get_initial_population_condition, calc/datasets.py:143-173: hospital figures of the start date from the area's
    case file, the unmeasurable ones from the variables; a start date the file does not hold gives the empty
    condition (the reference prints a note and does the same, :151-156)."""
    area = area or variables['area_name']
    row = _inputs().get('cases', {}).get(area, {}).get(variables['start_date'])
    if row is None:
        return InitialPopulationCondition()
    dead, in_icu, in_ward, confirmed = row
    return InitialPopulationCondition(
        dead=dead, in_icu=in_icu, in_ward=in_ward, confirmed_cases=confirmed,
        ill=variables.get('ill_at_simulation_start', 0), incubating=variables.get('incubating_at_simulation_start', 0),
        recovered=variables.get('recovered_at_simulation_start', 0))


def build_context_args(variables=None, area=None, age_count_override=None):
    """Arguments of model.Context as simulate_individuals builds them, with numpy inputs.

    `age_structure` is a dict-like {age: count}; `contacts_per_day` the long list from
    contacts_long().  reina_b200.model.Context accepts these as well as the pandas objects the
    real calc.simulation passes."""
    v = variables or default_variables()
    area = area or v['area_name']
    counts = age_counts(area) if age_count_override is None else np.asarray(age_count_override)
    pop_params = dict(
        age_structure={int(a): int(n) for a, n in enumerate(counts)},
        contacts_per_day=contacts_long(v['max_age']),
        initial_population_condition=initial_population_condition(v, area),
        age_groups=make_age_groups(v['max_age']),
        imported_infection_ages=v['imported_infection_ages'],
    )
    hc_params = dict(hospital_beds=v['hospital_beds'], icu_units=v['icu_units'])
    return dict(population_params=pop_params, healthcare_params=hc_params,
                disease_params=create_disease_params(v), start_date=v['start_date'],
                random_seed=v['random_seed'])


def active_interventions(variables=None, scenario=None):
    v = variables or default_variables()
    names = tuple(x['name'] for x in v.get('variants', []))
    tuples = v['interventions'] if scenario is None else scenario_interventions(scenario, v['interventions'])
    return [iv_tuple_to_obj(iv, names) for iv in tuples]
