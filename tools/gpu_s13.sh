#!/bin/bash
O=gpurun_out/s13; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/parity.log 2>&1; echo "parity rc=$?"
tail -4 $O/parity.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -q -m gpu -x -k "bit_exact" > $O/full.log 2>&1; echo "full rc=$?"
tail -3 $O/full.log
cat > /tmp/single.py <<'P'
import os, sys, numpy as np
sys.path.insert(0, '.')
import bench
R = int(sys.argv[1])
ctx = bench.make_context(bench.workload_spec('hus'), R, 0, 180, seed=1)
ms = []
for s in range(6):
    ctx.reset(60 + s); ctx.run(180)
    if s >= 2: ms.append(ctx._engine.last_step_ms())
print('R=%d RB_PDL=%s: %.3f ms per 180 days = %.1f us/day  %.3e agent-days/s' % (R, os.environ.get('RB_PDL'), np.mean(ms), np.mean(ms) / 180 * 1e3, 1685983*180*R/np.mean(ms)*1e3), flush=True)
P
for R in 1 32 64 256; do for p in 0 1; do RB_PDL=$p python /tmp/single.py $R; done; done > $O/pdl.log 2>&1
cat $O/pdl.log
