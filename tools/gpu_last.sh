#!/bin/bash
# what the driver runs at round end, on the final tree: the GPU suite, smoke(), the default bench
O=gpurun_out/last; mkdir -p $O
timeout 240 python -m pytest tests -x -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open('gpurun_out/last/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['traffic'], d['roofline']['frac_dram'], d['roofline']['isolated'], d['roofline_whole_run']['frac'])
P
