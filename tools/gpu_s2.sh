#!/bin/bash
# session 2: whole GPU test suite, ncu captures of the per-day kernels, per-kernel times at R = 1 / 32 / 256, default bench
O=gpurun_out/s2; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -rs --durations=15 > $O/gpu_tests.log 2>&1; echo "tests rc=$?"
tail -25 $O/gpu_tests.log
for R in 1 32 256; do python tools/kern_times.py $R; done > $O/kern_times.log 2>&1
cat $O/kern_times.log
for k in k_sweep k_expose k_resolve k_between; do
  RB_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 92 -c 1 -f \
      -o $O/day92_${k}_R256 python tools/prof_run.py --replicas 256 --days 95 > $O/full92_$k.log 2>&1
  python tools/ncu_report.py $O/day92_${k}_R256.ncu-rep 24 > $O/day92_${k}_R256.txt 2>&1
done
RB_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 45 -c 1 -f \
    -o $O/day45_k_sweep_R256 python tools/prof_run.py --replicas 256 --days 47 > $O/full45.log 2>&1
python tools/ncu_report.py $O/day45_k_sweep_R256.ncu-rep 24 > $O/day45_k_sweep_R256.txt 2>&1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -c 1500 $O/bench.json; tail -5 $O/bench.err
ls -la $O
