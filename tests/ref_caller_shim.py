"""Run the reference's OWN caller -- calc/simulation.py:148-290 `simulate_individuals`, unmodified, imported from
$REF (default /root/reference) -- over a chosen engine module bound as `cythonsim.model`:

    python tests/ref_caller_shim.py {ref|shim-oracle|shim-cuda} OUT.pkl [--days D] [--area A] [--seed S]

  ref          the unmodified Cython engine built in oracle/_ref (what the reference itself runs)
  shim-oracle  reina_b200.model (the drop-in module this repo ships) with the CPU oracle library injected through the
               test-only `_library` seam -- exercises the whole Python surface the caller touches without a GPU
  shim-cuda    reina_b200.model as shipped (CUDA); needs a GPU and the reference tree on the same machine

TEST INFRASTRUCTURE.  The reference application needs flask / flask_babel / flask_caching (sessions, i18n, cache) and
xlrd (hospital-district spreadsheet), none of which exist in this image; they are replaced by the minimal stand-ins
below, exactly the set SURVEY.md section 8c lists.  Nothing of the reference is copied: its modules are imported from
where they lie.  Runs in its own process so that the sys.modules surgery never leaks into pytest.
"""
import argparse
import os
import pickle
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('REF', '/root/reference')


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stand_ins():
    class _Session(dict):
        pass
    _module('flask', has_request_context=lambda: False, session=_Session())
    _module('flask_babel', lazy_gettext=lambda s, **kw: s, gettext=lambda s, **kw: s)

    class SimpleCache:
        def __init__(self, *a, **kw):
            self.d = {}

        def get(self, k):
            return self.d.get(k)

        def set(self, k, v, timeout=None):
            self.d[k] = v

    class Cache(SimpleCache):
        def init_app(self, app):
            pass

        def memoize(self, *a, **kw):
            return lambda f: f
    fc = _module('flask_caching', Cache=Cache)
    fc.backends = _module('flask_caching.backends')
    fc.backends.simple = _module('flask_caching.backends.simple', SimpleCache=SimpleCache)


def bind_engine(which):
    """sys.modules['cythonsim'] = a package whose `.model` is the chosen engine module."""
    if which == 'ref':
        sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
        # main.pyx:13 imports faker's Finnish name provider (debug names only): the stand-in under oracle/stubs
        import importlib.util
        for name, rel in (('faker', 'faker/__init__.py'), ('faker.providers', 'faker/providers/__init__.py'),
                          ('faker.providers.person', 'faker/providers/person/__init__.py'),
                          ('faker.providers.person.fi_FI', 'faker/providers/person/fi_FI.py')):
            spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'oracle', 'stubs', rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
        import cythonsim                       # oracle/_ref/cythonsim: the two compiled reference modules
        from cythonsim import main
        cythonsim.model = main
        return main
    sys.path.insert(0, ROOT)
    from reina_b200 import model
    if which == 'shim-oracle':
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import helpers
        lib = helpers.oracle_library()
        real = model.Context

        class Context(real):                   # same class, engine library injected (tests only)
            def __init__(self, *a, **kw):
                kw.setdefault('_library', lib)
                real.__init__(self, *a, **kw)
        model.Context = Context
    pkg = _module('cythonsim', model=model)
    pkg.__path__ = []
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('engine', choices=['ref', 'shim-oracle', 'shim-cuda'])
    ap.add_argument('out')
    ap.add_argument('--days', type=int, default=30)
    ap.add_argument('--area', default='Varsinais-Suomi')
    ap.add_argument('--seed', type=int, default=3)
    a = ap.parse_args()

    install_stand_ins()
    sys.path.insert(0, REF)                    # before bind_engine: oracle/_ref/cythonsim must shadow $REF/cythonsim (pyximport)
    # calc/datasets.py resolves its dataset directory at import time (and would mkdir inside the read-only reference tree)
    os.environ.setdefault('DATASET_PATH', '/tmp/reina_ref_datasets')
    os.makedirs(os.environ['DATASET_PATH'], exist_ok=True)
    open(os.path.join(os.environ['DATASET_PATH'], 'hosp_cases_turku.csv'), 'a').close()   # a filedep that is only stat()ed
    engine = bind_engine(a.engine)

    import pandas as pd
    import variables
    import calc.datasets
    import common.interventions

    # hospital-district membership comes from an .xls (xlrd is absent): the same municipality lists, recorded in
    # tools/make_inputs.py from that spreadsheet
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import make_inputs
    rows = [(m, hcd) for hcd, ms in make_inputs.MUNICIPALITIES.items() for m in ms]
    hdf = pd.DataFrame(rows, columns=['kunta', 'sairaanhoitopiiri'])
    calc.datasets.get_healthcare_districts = lambda **kw: hdf
    # interventions generated from downloaded Google-mobility / THL files: empty for every benchmark config
    common.interventions.generate_mobility_ivs = lambda **kw: []
    common.interventions.generate_vaccination_ivs = lambda **kw: []

    import calc.simulation as sim
    assert sim.model is engine, 'calc.simulation did not bind the chosen engine'
    with variables.allow_set_variable():
        variables.set_variable('simulation_days', a.days)
        variables.set_variable('area_name', a.area)
        variables.set_variable('random_seed', a.seed)
        df, adf = sim.simulate_individuals(skip_cache=True)
    with open(a.out, 'wb') as f:
        pickle.dump(dict(df=df, adf=adf, engine=a.engine, engine_file=engine.__file__), f)
    print('ok', a.engine, df.shape, adf.shape)


if __name__ == '__main__':
    main()
