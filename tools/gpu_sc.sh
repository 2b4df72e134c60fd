#!/bin/bash
# bench lines of the two remaining scenario lists of BASELINE configs[2]
mkdir -p gpurun_out/sc
for s in summer-boogie looser-restrictions-to-start-with; do
  timeout 120 python bench.py --workload scenario:$s --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/sc/bench_scenario_$s.json 2>> gpurun_out/sc/err.log; echo "$s rc=$?"
done
