#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): ncu evidence for profiles/.  Output goes to gpurun_out/prof/.
#   tools/collect_profiles.sh [R]
R=${1:-256}
O=gpurun_out/prof
mkdir -p $O
# 1. launch list of one 180-day run (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_R${R}.csv \
    python tools/prof_run.py --replicas $R --days 180 > $O/launches.log 2>&1
# 2. DRAM traffic of every k_sweep launch of the run (mean = roofline.traffic of bench.py).  One replica group, so
#    that one launch = one day of all R replicas, the unit bench.py's algorithmic bytes are quoted for.
RB_GROUPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_sweep --csv \
    --log-file $O/sweep_traffic_R${R}.csv python tools/prof_run.py --replicas $R --days 180 > $O/traffic.log 2>&1
# 3. full captures (one replica group: launch i of a kernel = day i): peak day (92) of every kernel, sparse day (45) of the sweep
for k in k_sweep k_expose k_resolve k_between; do
  RB_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:$k -s 92 -c 1 -f \
      -o $O/day92_${k}_R${R} python tools/prof_run.py --replicas $R --days 95 > $O/full92_$k.log 2>&1
done
RB_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 45 -c 1 -f \
    -o $O/day45_k_sweep_R${R} python tools/prof_run.py --replicas $R --days 47 > $O/full45.log 2>&1
ls -la $O
