#!/bin/bash
# session 1: does the list-based sweep agree with the oracle; first timings
O=gpurun_out/s1; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" 
tail -3 $O/smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/parity.log 2>&1; echo "parity rc=$?"
tail -15 $O/parity.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_quick.json 2> $O/bench_quick.err; echo "bench rc=$?"
tail -c 3000 $O/bench_quick.json; tail -5 $O/bench_quick.err
