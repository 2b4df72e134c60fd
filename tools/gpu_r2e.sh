#!/bin/bash
# round 2, session 2, call e: stage T of the sweep as its own kernel (k_sweep<true> + k_slow): parity, step time, occupancy variants
O=gpurun_out/e1; mkdir -p $O
RB_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $O/parity_split.log 2>&1; echo "parity(split forced) rc=$?"; tail -3 $O/parity_split.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -x -q -k "bench_geometry" > $O/geometry.log 2>&1; echo "geometry rc=$?"; tail -3 $O/geometry.log
{
echo "== fused (RB_SPLIT=0)"; RB_SPLIT=0 timeout 300 python tools/group_exp.py --replicas 256 --configs 4:50 --steps 4
echo "== split, 12 CTAs/SM (default)"; timeout 300 python tools/group_exp.py --replicas 256 --configs 4:50,2:100,4:100 --steps 4
for v in sp10 sp14 sp16; do echo "== split $v"; REINA_B200_LIB=build/variants/$v.so timeout 300 python tools/group_exp.py --replicas 256 --configs 4:50 --steps 4; done
echo "== R=32 fused / split"; RB_SPLIT=0 timeout 300 python tools/group_exp.py --replicas 32 --configs 4:50 --steps 4; timeout 300 python tools/group_exp.py --replicas 32 --configs 4:50 --steps 4
python tools/kern_times.py 256
} 2>&1 | tee $O/split_experiment.txt
