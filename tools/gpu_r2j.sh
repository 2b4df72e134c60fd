#!/bin/bash
# round 2, session 2, call j (8 GPUs): the population-sharded tests on 1 / 2 / 4 / 8 real ranks with the final code (graph-replayed days)
O=gpurun_out/s8; mkdir -p $O
nvidia-smi -L | wc -l
timeout 560 python -m pytest tests/test_gpu_sharded.py -v -rs > $O/sharded_tests_8gpu.log 2>&1; echo "sharded rc=$?"
grep -E "PASS|FAIL|ERROR|SKIP|passed|failed" $O/sharded_tests_8gpu.log | tail -22
