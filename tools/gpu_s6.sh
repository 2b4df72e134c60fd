#!/bin/bash
# session 6: one-wave k_resolve; replica-group experiments
O=gpurun_out/s6; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/parity.log 2>&1; echo "parity rc=$?"
tail -3 $O/parity.log
for rs in 2 4 8 16; do RB_RESOLVE_PER_SM=$rs python tools/kern_times.py 256; done > $O/kern_times.log 2>&1
for R in 1 32; do python tools/kern_times.py $R; done >> $O/kern_times.log 2>&1
cat $O/kern_times.log
python tools/group_exp.py --replicas 256 --configs 4:50,4:100,6:50,8:50,8:34,8:25 --steps 3 > $O/group_exp.log 2>&1
RB_RESOLVE_PER_SM=8 python tools/group_exp.py --replicas 256 --configs 4:50,8:50 --steps 3 >> $O/group_exp.log 2>&1
python tools/group_exp.py --replicas 32 --configs 1:100,2:100,4:100,8:100,8:50,4:50 --steps 3 >> $O/group_exp.log 2>&1
cat $O/group_exp.log
