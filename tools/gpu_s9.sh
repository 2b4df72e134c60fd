#!/bin/bash
# session 9: the persistent run kernel (few replicas)
O=gpurun_out/s9; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/parity.log 2>&1; echo "parity rc=$?"
tail -15 $O/parity.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -q -m gpu -x -k "bit_exact or properties or worker or monte" > $O/full.log 2>&1; echo "full rc=$?"
tail -5 $O/full.log
cat > /tmp/single.py <<'P'
import os, sys, numpy as np
sys.path.insert(0, '.')
import bench
R = int(sys.argv[1])
ctx = bench.make_context(bench.workload_spec('hus'), R, 0, 180, seed=1)
ms = []
for s in range(5):
    ctx.reset(60 + s); ctx.run(180)
    if s >= 2: ms.append(ctx._engine.last_step_ms())
print('R=%d PERSISTENT=%s RUN_CTAS=%s: %.3f ms per 180 days = %.1f us/day, %.3e agent-days/s' % (R, os.environ.get('RB_PERSISTENT'), os.environ.get('RB_RUN_CTAS'), np.mean(ms), np.mean(ms) / 180 * 1e3, 1685983 * 180 * R / np.mean(ms) * 1e3), flush=True)
P
for cfg in "0 -" "1 8" "1 16" "1 32" "1 64" "1 148" "1 296"; do set -- $cfg; RB_PERSISTENT=$1 RB_RUN_CTAS=$2 python /tmp/single.py 1; done > $O/single.log 2>&1
for cfg in "0 -" "1 2" "1 4" "1 9"; do set -- $cfg; RB_PERSISTENT=$1 RB_RUN_CTAS=$2 python /tmp/single.py 32; done >> $O/single.log 2>&1
for cfg in "0 -" "1 4" "1 2"; do set -- $cfg; RB_PERSISTENT=$1 RB_RUN_CTAS=$2 python /tmp/single.py 64; done >> $O/single.log 2>&1
cat $O/single.log
