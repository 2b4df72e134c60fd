#!/bin/bash
# round 2, session 2, call d: Philox2x32 for the 64-bit draws: parity + step time
O=gpurun_out/d1; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $O/parity.log 2>&1; echo "parity rc=$?"; tail -3 $O/parity.log
timeout 300 python tools/group_exp.py --replicas 256 --configs 4:50 --steps 4 2>&1 | tee $O/step.txt
timeout 300 python tools/group_exp.py --replicas 32 --configs 4:50 --steps 4 2>&1 | tee -a $O/step.txt
python tools/kern_times.py 256 2>&1 | tee -a $O/step.txt
