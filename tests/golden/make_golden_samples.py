"""Golden sampler histograms from the UNMODIFIED reference engine (oracle/_ref): Context.sample() for the two
kinds that work under Cython 3.3 (SURVEY.md section 8c "Known oracle defect"): 'contacts_per_day' and
'symptom_severity', 40 x 10000 draws per (kind, age).  Duration samplers are pinned against numpy's
Generator(PCG64).standard_gamma(float32) restated from simrandom.pyx:46-55 + main.pyx:977-1039 in the test itself.

    python tests/golden/make_golden_samples.py   ->  tests/golden/ref_samples.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

AGES = [5, 15, 25, 45, 65, 75, 85]


def main():
    ctx = ref_harness.make_context(area='Varsinais-Suomi', seed=123)
    out = {}
    for age in AGES:
        c = np.zeros(101, dtype=np.int64)
        s = np.zeros(5, dtype=np.int64)
        for _ in range(40):
            c += np.bincount(ctx.sample('contacts_per_day', age), minlength=101)[:101]
            s += np.bincount(ctx.sample('symptom_severity', age), minlength=5)[:5]
        out['contacts_%d' % age] = c
        out['severity_%d' % age] = s
    np.savez_compressed(os.path.join(HERE, 'ref_samples.npz'), ages=np.array(AGES), **out)
    print({k: v[:8] for k, v in out.items()})


if __name__ == '__main__':
    main()
