"""One-off check of the FIRST 45 days at high power (test infrastructure; needs oracle/_ref): 4096 oracle seeds against 4096
seeds of the unmodified reference engine, seed sets disjoint from every other comparison.  Run after the 1024-seed checks
showed one early cell (deaths on day 32, ~9 per run) at z = -3.2 in all three scenarios -- which share their seeds and their
first weeks -- to tell a fluctuation of that seed set from a timing difference.  Output: early_check_hus_default_4096.txt."""
import os, sys, time
import numpy as np
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests')
import helpers
from oracle import ref_harness
DAYS=45; N=4096
def _oracle(seed):
    ctx = helpers.make_context(helpers.oracle_library(), area='HUS', seed=seed, max_days=DAYS+1)
    ctx.run(DAYS)
    return helpers.series_matrix(ctx)[0]
if __name__ == '__main__':
    t0=time.time()
    ref,_,_ = ref_harness.run_ensemble(np.arange(200000, 200000+N), days=DAYS, area='HUS', scenario=None)
    t1=time.time()
    with ProcessPoolExecutor(8) as ex:
        mine = np.stack(list(ex.map(_oracle, [300000+s for s in range(N)], chunksize=16)))
    t2=time.time()
    names = helpers.series_names()
    se = np.sqrt(ref.std(0, ddof=1)**2/N + mine.std(0, ddof=1)**2/N)
    diff = mine.mean(0)-ref.mean(0)
    with np.errstate(divide='ignore', invalid='ignore'):
        z = np.where(se>0, diff/se, 0.0)
    print('reference %.0f s, oracle %.0f s; cells beyond 3 SE %.4f, beyond 2 SE %.4f, worst |z| %.2f' % (t1-t0, t2-t1, (np.abs(z)>3).mean(), (np.abs(z)>2).mean(), np.abs(z).max()))
    for s in ('all_infected','dead','all_detected','recovered','in_ward','in_icu','infected'):
        j=names.index(s)
        print(s, ' '.join('d%d:%+.2f(%.1f/%.1f)' % (d, z[d,j], mine.mean(0)[d,j], ref.mean(0)[d,j]) for d in (5,10,15,20,25,28,30,31,32,33,35,40,44)))
    