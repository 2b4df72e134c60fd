"""One-off, higher-power statistical comparison of the CPU oracle with the UNMODIFIED reference engine (test
infrastructure; needs oracle/_ref, i.e. the build container):
    python tests/golden/power_check.py [--seeds 256] [--area HUS] [--scenario ...]
N reference seeds against N oracle seeds at full size, every daily series; prints the fraction of (day, series) cells
beyond 3 standard errors, the worst cells and the day-180 totals.  The committed goldens hold 64 reference seeds
(256 for Varsinais-Suomi); this script is how the figures quoted in DESIGN.md section 2 were obtained."""
import argparse
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import helpers  # noqa: E402
from oracle import ref_harness  # noqa: E402

A = None


def _oracle(seed):
    ctx = helpers.make_context(helpers.oracle_library(), area=A.area, scenario=A.scenario, seed=seed, max_days=A.days + 1)
    ctx.run(A.days)
    return helpers.series_matrix(ctx)[0]


def main():
    global A
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', type=int, default=256)
    ap.add_argument('--area', default='HUS')
    ap.add_argument('--scenario', default=None)
    ap.add_argument('--days', type=int, default=180)
    ap.add_argument('--out', default=None)
    A = ap.parse_args()
    t0 = time.time()
    ref, _, _ = ref_harness.run_ensemble(np.arange(50000, 50000 + A.seeds), days=A.days, area=A.area, scenario=A.scenario)
    t1 = time.time()
    with ProcessPoolExecutor(os.cpu_count() or 1) as ex:
        mine = np.stack(list(ex.map(_oracle, [70000 + s for s in range(A.seeds)])))
    t2 = time.time()
    names = helpers.series_names()
    se = np.sqrt(ref.std(0, ddof=1) ** 2 / ref.shape[0] + mine.std(0, ddof=1) ** 2 / mine.shape[0])
    diff = mine.mean(0) - ref.mean(0)
    with np.errstate(divide='ignore', invalid='ignore'):
        z = np.where(se > 0, diff / se, 0.0)
    exact = (se == 0) & (np.abs(diff) > 1e-9)
    print('%s / %s, %d + %d seeds x %d days: reference %.0f s, oracle %.0f s' % (A.area, A.scenario, A.seeds, A.seeds, A.days, t1 - t0, t2 - t1))
    print('deterministic series that differ:', sorted({names[j] for j in np.argwhere(exact)[:, 1]}))
    print('cells beyond 3 SE: %.4f   beyond 2 SE: %.4f (a normal gives 0.0027 / 0.0455)   worst |z| %.2f' % (
        (np.abs(z) > 3).mean(), (np.abs(z) > 2).mean(), np.abs(z).max()))
    order = np.dstack(np.unravel_index(np.argsort(-np.abs(z), axis=None)[:8], z.shape))[0]
    for d, j in order:
        print('   day %3d %-28s z %+.2f   oracle %.1f   reference %.1f' % (d, names[j], z[d, j], mine.mean(0)[d, j], ref.mean(0)[d, j]))
    for s in ('all_infected', 'dead', 'all_detected', 'recovered', 'cum_icu', 'in_ward', 'in_icu', 'exposed_per_day'):
        j = names.index(s)
        rel = diff[-1, j] / ref.mean(0)[-1, j] if ref.mean(0)[-1, j] else 0.0
        print('   last day %-18s oracle %12.1f  reference %12.1f  (%+.2f %%, z %+.2f)' % (s, mine.mean(0)[-1, j], ref.mean(0)[-1, j], 100 * rel, z[-1, j]))
    if A.out:
        np.savez_compressed(A.out, z=z, names=np.array(names), ref_mean=ref.mean(0), oracle_mean=mine.mean(0))


if __name__ == '__main__':
    main()
