#!/bin/bash
O=gpurun_out/s10; mkdir -p $O
for c in 16 64 148; do RB_PERSISTENT=1 RB_RUN_CTAS=$c python tools/run_phases.py 1; done > $O/phases.log 2>&1
RB_PERSISTENT=1 RB_RUN_CTAS=9 python tools/run_phases.py 32 >> $O/phases.log 2>&1
cat $O/phases.log
