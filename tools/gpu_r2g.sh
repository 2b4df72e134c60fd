#!/bin/bash
# round 2, session 2, call g: smaller day-boundary CTAs (easier to place between another group's sweep CTAs)
O=gpurun_out/h1; mkdir -p $O
for v in pre256 pre128; do
  echo "== $v"
  REINA_B200_LIB=build/variants/$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
  REINA_B200_LIB=build/variants/$v.so timeout 300 python tools/group_exp.py --replicas 256 --configs 4:50 --steps 3
  REINA_B200_LIB=build/variants/$v.so timeout 300 python tools/group_exp.py --replicas 32 --configs 4:50 --steps 3
  REINA_B200_LIB=build/variants/$v.so timeout 300 python tools/kern_times.py 1
done 2>&1 | tee $O/pre_threads.txt
