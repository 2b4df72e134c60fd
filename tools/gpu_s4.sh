#!/bin/bash
# session 4: two-ended segments, pipelined item reservation, device-built 4096-cell row guide
O=gpurun_out/s5; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q -m gpu -x > $O/parity.log 2>&1; echo "parity rc=$?"
tail -5 $O/parity.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -q -m gpu -x -k "bit_exact or properties" > $O/full.log 2>&1; echo "full rc=$?"
tail -5 $O/full.log
python tools/kern_times.py 256 > $O/kern_times.log 2>&1
for v in sw8 sw10 ex12 ex36 ex48 exr10; do REINA_B200_LIB=build/variants/$v.so python tools/kern_times.py 256 >> $O/kern_times.log 2>&1; done
cat $O/kern_times.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > $O/bench_quick.json 2> $O/bench_quick.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/s5/bench_quick.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline_whole_run']['frac'], d['kernel_ms_per_day'])
P
for k in k_sweep k_expose; do
  RB_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 92 -c 1 -f \
      -o $O/day92_${k}_R256 python tools/prof_run.py --replicas 256 --days 95 > $O/full92_$k.log 2>&1
  python tools/ncu_report.py $O/day92_${k}_R256.ncu-rep 30 > $O/day92_${k}_R256.txt 2>&1
done
