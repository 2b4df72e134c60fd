#!/bin/bash
O=gpurun_out/s12; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q -m gpu -x > $O/parity.log 2>&1; echo "parity rc=$?"
tail -4 $O/parity.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -q -m gpu -x -k "bit_exact" > $O/full.log 2>&1; echo "full rc=$?"
tail -3 $O/full.log
python tools/kern_times.py 256 > $O/kern_times.log 2>&1; cat $O/kern_times.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_quick.json 2> $O/bench_quick.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/s12/bench_quick.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline_whole_run']['frac'], d['kernel_ms_per_day']['isolated'], d['roofline']['traffic'], d['roofline']['traffic_source'])
P
