#!/bin/bash
# round 2, session 2, call a: local ranks + sharded graphs on one GPU, the high-power statistics test, launch priorities
O=gpurun_out/a1; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharded.py -v -rs > $O/sharded_1gpu.log 2>&1; echo "sharded rc=$?"
grep -E "PASS|FAIL|ERROR|SKIP|passed|failed" $O/sharded_1gpu.log | tail -25
timeout 400 python -m pytest tests/test_gpu_full_size.py -v -s -k "high_power" > $O/high_power.log 2>&1; echo "hp rc=$?"
grep -E "hus_default|day 180|passed|failed|Error" $O/high_power.log | tail -12
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
print('R=%d RB_PRIO=%s: %.3f ms per 180 days = %.1f us/day' % (R, os.environ.get('RB_PRIO'), np.mean(ms), np.mean(ms) / 180 * 1e3), flush=True)
P
for prio in 0 1; do
  echo "== RB_PRIO=$prio"
  RB_PRIO=$prio timeout 300 python tools/group_exp.py --replicas 256 --configs 4:50,2:100 --steps 3
  RB_PRIO=$prio timeout 300 python tools/group_exp.py --replicas 32 --configs 2:100,4:50,4:100 --steps 3
  RB_PRIO=$prio timeout 120 python /tmp/single.py 1
done 2>&1 | tee $O/priority_experiment.txt
echo "== sw10"
REINA_B200_LIB=build/variants/sw10.so timeout 300 python tools/group_exp.py --replicas 256 --configs 4:50 --steps 3 2>&1 | tee $O/sw10.txt
