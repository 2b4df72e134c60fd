#!/bin/bash
O=gpurun_out/s14; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q -m gpu -x > $O/parity.log 2>&1; echo "parity rc=$?"
tail -4 $O/parity.log
timeout 900 python -m pytest tests/test_gpu_full_size.py -q -m gpu -x > $O/full.log 2>&1; echo "full rc=$?"
tail -4 $O/full.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_quick.json 2> $O/bench_quick.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/s14/bench_quick.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline_whole_run']['frac'], d['gpu_launches'])
P
