"""Pre-check of tests/test_gpu_full_size.py::test_high_power_statistics WITHOUT a GPU (test infrastructure):
    python tests/golden/precheck_high_power.py [golden names ...]
The CUDA engine is bit-identical to the sequential oracle, so the 256 replicas the GPU test runs (replica r = seed + r)
can be reproduced here with the oracle, one process per seed, and pushed through the very assertions of the test."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import helpers  # noqa: E402


def _run(job):
    import test_gpu_full_size as T
    area, scenario, seed, variables = job
    ctx = helpers.make_context(helpers.oracle_library(), area=area, scenario=scenario, seed=seed, max_days=181,
                               variables=T.high_power_variables(variables))
    ctx.run(180)
    return helpers.series_matrix(ctx)[0]


def main():
    import test_gpu_full_size as T
    want = set(sys.argv[1:])
    for gold_name, area, scenario, seed, variables in T.HIGH_POWER:
        if want and gold_name not in want:
            continue
        with ProcessPoolExecutor(os.cpu_count() or 1) as ex:
            mine = np.stack(list(ex.map(_run, [(area, scenario, seed + r, variables) for r in range(256)])))
        T.high_power_report(mine, gold_name)
        print('%s: the assertions of the GPU test hold' % gold_name, flush=True)


if __name__ == '__main__':
    main()
