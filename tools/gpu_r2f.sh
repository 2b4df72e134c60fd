#!/bin/bash
# round 2, session 2, call f: finer scan of replica groups x wave share (256 replicas)
O=gpurun_out/g1; mkdir -p $O
timeout 600 python tools/group_exp.py --replicas 256 --configs 4:50,3:67,3:50,4:40,4:60,4:67,5:40,5:50,6:34,6:40,6:50 --steps 3 2>&1 | tee $O/group_scan.txt
for b in 26 104; do echo "== RB_RESOLVE_BLOCKS=$b"; RB_RESOLVE_BLOCKS=$b timeout 300 python tools/group_exp.py --replicas 256 --configs 4:50 --steps 3; done 2>&1 | tee -a $O/group_scan.txt
