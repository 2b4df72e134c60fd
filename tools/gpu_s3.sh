#!/bin/bash
# session 3: segmented lists + one-load row pick: parity, then kernel times of the default build and tuning variants
O=gpurun_out/s3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q -m gpu -x > $O/parity.log 2>&1; echo "parity rc=$?"
tail -5 $O/parity.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -q -m gpu -x -k "bit_exact or properties" > $O/full.log 2>&1; echo "full rc=$?"
tail -5 $O/full.log
python tools/kern_times.py 256 > $O/kern_times.log 2>&1
for v in sw8 sw9 sw12 g9 g11 ex10 ex16; do REINA_B200_LIB=build/variants/$v.so python tools/kern_times.py 256 >> $O/kern_times.log 2>&1; done
cat $O/kern_times.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > $O/bench_quick.json 2> $O/bench_quick.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/s3/bench_quick.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline_whole_run']['frac'], d['kernel_ms_per_day'])
P
for cfg in "32 1:100,2:100,4:100,8:100,8:50" "256 4:50,4:100,8:50,2:100"; do set -- $cfg; python tools/group_exp.py --replicas $1 --configs $2 --steps 3; done > $O/group_exp.log 2>&1
cat $O/group_exp.log
