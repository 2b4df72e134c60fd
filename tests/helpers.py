"""Shared test helpers: bind the CPU oracle / the CUDA library through the same ctypes wrapper."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from reina_b200 import _abi, inputs, model  # noqa: E402

ORACLE_PATH = os.path.join(ROOT, 'oracle', 'libreina_oracle.so')
_oracle = None

# order of oracle/ref_harness.series_names()
POP_SERIES = ['susceptible', 'vaccinated', 'infected', 'all_infected', 'detected', 'all_detected',
              'in_icu', 'cum_icu', 'in_ward', 'dead', 'recovered', 'non_hospital_deaths', 'new_infections']
SCALAR_SERIES = ['available_icu_units', 'available_hospital_beds', 'total_icu_units', 'r',
                 'exposed_per_day', 'ct_cases_per_day', 'mobility_limitation']
PLACES = ['home', 'work', 'school', 'transport', 'leisure', 'other']


def oracle_library():
    global _oracle
    if _oracle is None:
        subprocess.run(['make', '-s', '-C', os.path.join(ROOT, 'oracle')], check=True)
        _oracle = _abi.Library(ORACLE_PATH, 'ro_')
    return _oracle


def cuda_library():
    return _abi.cuda_library()


def make_context(lib, area='HUS', scenario=None, seed=0, n_replicas=1, variables=None,
                 age_count_override=None, interventions=None, max_days=200, **kw):
    v = variables or inputs.default_variables()
    args = inputs.build_context_args(v, area=area, age_count_override=age_count_override)
    args['random_seed'] = seed
    ctx = model.Context(n_replicas=n_replicas, max_days=max_days, _library=lib, **args, **kw)
    ivs = interventions if interventions is not None else inputs.active_interventions(v, scenario)
    for iv in ivs:
        ctx.add_intervention(iv)
    return ctx


def small_population(total, area='HUS'):
    return inputs.synthetic_age_counts(total, area)


def series_matrix(ctx, days=None):
    """[replica, day, series] float64 in the order of oracle/ref_harness.series_names()."""
    rows = ctx.series(0, days)
    R, D, _ = rows.shape
    G = len(ctx.age_group_labels)
    nA = len(_abi.ATTRS)
    out = np.zeros((R, D, len(POP_SERIES) + len(SCALAR_SERIES) + len(PLACES) + len(ctx.variant_names)))
    sc = rows[:, :, nA * G:].astype(np.float64)
    S = {name: sc[:, :, i] for i, name in enumerate(_abi.SCALARS)}
    k = 0
    for a in POP_SERIES:
        i = _abi.ATTRS.index(a)
        out[:, :, k] = rows[:, :, i * G:(i + 1) * G].sum(axis=2)
        k += 1
    with np.errstate(divide='ignore', invalid='ignore'):
        r = np.where(S['total_infectors'] > 5, S['total_infections'] / np.maximum(S['total_infectors'], 1), 0.0)
    mob = np.vectorize(lambda ep: ctx._epoch_mobility.get(int(ep), 0.0))(S['table_epoch'])
    for name in SCALAR_SERIES:
        out[:, :, k] = r if name == 'r' else (mob if name == 'mobility_limitation' else S[name])
        k += 1
    place_idx = {'home': 0, 'work': 1, 'school': 2, 'transport': 3, 'leisure': 4, 'other': 5}
    for p in PLACES:
        out[:, :, k] = sc[:, :, _abi.RB_S_CONTACTS0 + place_idx[p]]
        k += 1
    for i in range(len(ctx.variant_names)):
        out[:, :, k] = sc[:, :, _abi.RB_S_VARIANT0 + i]
        k += 1
    return out


def series_names(variant_names=('wild-type', 'b1.1.7')):
    return (POP_SERIES + SCALAR_SERIES + ['exposures_%s' % p for p in PLACES]
            + ['infected_by_variant_%s' % v for v in variant_names])


# ---------------------------------------------------------------------------------------------------
# parity utilities
# ---------------------------------------------------------------------------------------------------
def scaled_capacity(n_agents, beds=2600, icu=300, full=1685983):
    """Hospital capacity scaled with the population so that small test populations still saturate."""
    return max(1, round(beds * n_agents / full)), max(1, round(icu * n_agents / full))


def stress_interventions():
    """A schedule that exercises every intervention type early: imports of both variants, weekly trickle
    with a variant share, every testing mode, contact tracing, masks, place/age mobility limits,
    vaccination programmes and capacity building."""
    T = [
        ['import-infections', '2020-02-18', 150],
        ['import-infections', '2020-02-18', 60, 'b1.1.7'],
        ['test-all-with-symptoms', '2020-02-20'],
        ['import-infections-weekly', '2020-02-21', 40, 30],
        ['limit-mobility', '2020-02-25', 20],
        ['test-only-severe-symptoms', '2020-02-27', 40],
        ['wear-masks', '2020-02-28', 60, 15, None, None],
        ['limit-mobility', '2020-03-01', 50, 7, 18, 'school'],
        ['test-with-contact-tracing', '2020-03-04', 60],
        ['vaccinate', '2020-03-05', 700, 70, None],
        ['vaccinate', '2020-03-08', 1400, 16, 69],
        ['import-infections', '2020-03-10', 100],
        ['limit-mobility', '2020-03-12', 0],
        ['wear-masks', '2020-03-12', 90, None, None, 'transport'],
        ['build-new-hospital-beds', '2020-03-20', 3],
        ['build-new-icu-units', '2020-03-20', 1],
        ['test-with-contact-tracing', '2020-03-25', 100],
        ['vaccinate', '2020-03-28', 2100, 70, None],
        ['test-all-with-symptoms', '2020-04-10'],
        ['test-with-contact-tracing', '2020-04-20', 35],
    ]
    from reina_b200 import inputs as _inputs
    return [_inputs.iv_tuple_to_obj(t) for t in T]


def diff_report(gpu, cpu, days):
    """Compare two contexts after the same run; returns a list of human-readable differences (empty = equal)."""
    out = []
    a, b = gpu.series(0, days), cpu.series(0, days)
    names = gpu.row_layout()
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        r, d, col = bad[0]
        cols = sorted({names[c] for c in bad[bad[:, 1] == d][:, 2]})
        out.append('stats differ at %d cells; first: replica %d day %d col %s gpu=%d cpu=%d; columns that day: %s'
                   % (len(bad), r, d, names[col], a[r, d, col], b[r, d, col], cols[:12]))
    for r in range(gpu.n_replicas):
        ga, ca = gpu._engine.read_agents(r), cpu._engine.read_agents(r)
        if not np.array_equal(ga, ca):
            for f in ga.dtype.names:
                if not np.array_equal(ga[f], ca[f]):
                    idx = np.flatnonzero(ga[f] != ca[f])
                    out.append('replica %d agent field %s differs for %d agents, first %d: gpu=%s cpu=%s'
                               % (r, f, len(idx), idx[0], ga[f][idx[0]], ca[f][idx[0]]))
        gq, cq = gpu._engine.read_queue(r), cpu._engine.read_queue(r)
        if sorted(gq.tolist()) != sorted(cq.tolist()):
            out.append('replica %d test queue differs: gpu %d entries, cpu %d' % (r, len(gq), len(cq)))
        if not np.array_equal(gpu._engine.read_available(r), cpu._engine.read_available(r)):
            out.append('replica %d free beds/icu differ' % r)
    return out


class TorchComm:
    """Adapter that gives a torch.distributed process group (gloo in the CPU tests) the communicator interface of
    reina_b200/comm.py: rank, size, allreduce(x, op), allgather(x), barrier()."""

    def __init__(self, dist):
        self.dist = dist
        self.rank, self.size = dist.get_rank(), dist.get_world_size()

    def allreduce(self, x, op='sum'):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == 'sum' else self.dist.ReduceOp.MAX)
        return t.numpy()

    def allgather(self, x):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(x).copy())
        out = [torch.empty_like(t) for _ in range(self.size)]
        self.dist.all_gather(out, t)
        return np.stack([o.numpy() for o in out])

    def barrier(self):
        self.dist.barrier()
