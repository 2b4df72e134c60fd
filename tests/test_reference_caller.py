"""The reference's real caller, calc/simulation.py:148-290 `simulate_individuals` (imported UNMODIFIED from
/root/reference by tests/ref_caller_shim.py, with stand-ins only for flask / flask_babel / flask_caching / xlrd),
driven over `reina_b200.model` bound as `cythonsim.model` -- the one-line integration INTEGRATION.md section 1 describes --
side by side with the same call over the unmodified Cython engine (oracle/_ref).

Needs the reference tree, so these run in the build container only (skipped on the GPU box, where /root/reference does
not exist); the engine under the shim is therefore the CPU oracle library, injected through the test-only `_library`
seam.  The CUDA engine runs behind exactly the same Python surface (tests/test_gpu_*.py compare the two bit for bit).
"""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

import helpers

REF = os.environ.get('REF', '/root/reference')
SHIM = os.path.join(helpers.ROOT, 'tests', 'ref_caller_shim.py')

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'calc')), reason='needs the reference tree')


def _run(engine, out, days, area='Varsinais-Suomi', seed=3):
    r = subprocess.run([sys.executable, SHIM, engine, str(out), '--days', str(days), '--area', area, '--seed', str(seed)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    with open(out, 'rb') as f:
        return pickle.load(f)


def test_real_caller_over_the_shim_matches_the_reference_engine(tmp_path):
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip('oracle/_ref not built')
    days = 30
    ref = _run('ref', tmp_path / 'ref.pkl', days)
    shim = _run('shim-oracle', tmp_path / 'shim.pkl', days)
    assert ref['engine_file'].startswith(os.path.join(helpers.ROOT, 'oracle', '_ref'))
    assert shim['engine_file'] == os.path.join(helpers.ROOT, 'reina_b200', 'model.py')
    a, b = ref['df'], shim['df']
    # identical frames as far as a caller can tell: columns, order, dtypes, index, shapes
    assert list(a.columns) == list(b.columns) and a.shape == b.shape == (days, 26)
    assert a.index.equals(b.index) and a.dtypes.equals(b.dtypes)
    aa, ab = ref['adf'], shim['adf']
    assert aa.columns.equals(ab.columns) and aa.index.equals(ab.index) and aa.dtypes.equals(ab.dtypes) and aa.shape == (days, 12 * 9)
    # deterministic cells agree exactly; the two engines use different random streams, so the rest is compared in
    # distribution by tests/test_oracle_pinned.py against >= 64 reference seeds
    assert a['susceptible'].iloc[0] == b['susceptible'].iloc[0] == 479861
    for col in ('total_icu_units', 'mobility_limitation', 'vaccinated'):
        assert np.array_equal(a[col].to_numpy(dtype=float), b[col].to_numpy(dtype=float)), col
    assert np.array_equal(aa['susceptible'].iloc[0].to_numpy(), ab['susceptible'].iloc[0].to_numpy())
    for f in (a, b):                                   # imports of 2020-02-22 .. 03-15 (default interventions) took hold in both
        assert f['all_infected'].iloc[-1] > 200 and f['susceptible'].iloc[-1] + f['all_infected'].iloc[-1] == 479861

    # the per-day API the real caller uses (generate_state + iterate, one day at a time) gives the same numbers as this
    # repo's batched caller (reina_b200/simulation.py: one run, one fetch) on the same engine and seed
    from reina_b200 import inputs, simulation
    v = inputs.default_variables(simulation_days=days, area_name='Varsinais-Suomi', random_seed=3)
    ctx = simulation.make_context(v, _library=helpers.oracle_library())
    df2, adf2 = simulation.simulate_individuals(v, context=ctx)
    cols = [c for c in b.columns if c != 'us_per_infected']
    assert np.array_equal(b[cols].to_numpy(dtype=float), df2[cols].to_numpy(dtype=float))
    assert np.array_equal(ab.to_numpy(), adf2.to_numpy()) and ab.columns.equals(adf2.columns)


@pytest.mark.gpu
def test_real_caller_over_the_cuda_shim(tmp_path):
    """Only where a GPU and the reference tree meet (never on the driver's boxes): the same call over the shipped module."""
    days = 30
    cuda = _run('shim-cuda', tmp_path / 'cuda.pkl', days)
    orac = _run('shim-oracle', tmp_path / 'shim.pkl', days)
    cols = [c for c in cuda['df'].columns if c != 'us_per_infected']
    assert np.array_equal(cuda['df'][cols].to_numpy(dtype=float), orac['df'][cols].to_numpy(dtype=float))
    assert np.array_equal(cuda['adf'].to_numpy(), orac['adf'].to_numpy())
