#!/bin/bash
# session 7: segments = 2 x production warps, resolve grid restored
O=gpurun_out/s7; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/parity.log 2>&1; echo "parity rc=$?"
tail -3 $O/parity.log
for R in 256 32 1; do python tools/kern_times.py $R; done > $O/kern_times.log 2>&1
for rb in 26 104 208; do RB_RESOLVE_BLOCKS=$rb python tools/kern_times.py 256; done >> $O/kern_times.log 2>&1
cat $O/kern_times.log
python tools/group_exp.py --replicas 256 --configs 4:50,4:100,8:50 --steps 3 > $O/group_exp.log 2>&1
python tools/group_exp.py --replicas 32 --configs 1:100,2:100,4:100,4:50,8:50 --steps 3 >> $O/group_exp.log 2>&1
cat $O/group_exp.log
