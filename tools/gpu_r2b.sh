#!/bin/bash
# round 2, session 2, call b: local ranks on one GPU (stop at the first failure), high-power statistics, priorities at 64 / 128 replicas
O=gpurun_out/b1; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -v -rs -k "local_ranks" > $O/sharded_local.log 2>&1; echo "local rc=$?"
grep -E "PASS|FAIL|ERROR|passed|failed|Error|error:|assert" $O/sharded_local.log | tail -25
timeout 300 python -m pytest tests/test_gpu_full_size.py -v -s -k "high_power" > $O/high_power.log 2>&1; echo "hp rc=$?"
grep -E "hus_default|day 180|passed|failed|Error" $O/high_power.log | tail -12
for R in 64 128; do for prio in 0 1; do
  echo "== R=$R RB_PRIO=$prio"
  RB_PRIO=$prio timeout 300 python tools/group_exp.py --replicas $R --configs 2:100,4:50 --steps 3
done; done 2>&1 | tee $O/priority_experiment_64_128.txt
