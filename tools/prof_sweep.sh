#!/bin/bash
# Runs ON THE GPU BOX: full ncu captures of k_sweep on a sparse day (45) and the peak day (92).  tools/prof_sweep.sh [R] [tag]
R=${1:-256}
TAG=${2:-x}
O=gpurun_out/prof
mkdir -p $O
export RB_GROUPS=1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 45 -c 1 -f \
    -o $O/sw45_${TAG} python tools/prof_run.py --replicas $R --days 47 > $O/sw45.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 92 -c 1 -f \
    -o $O/sw92_${TAG} python tools/prof_run.py --replicas $R --days 94 > $O/sw92.log 2>&1
ls -la $O
