#!/bin/bash
# round 2, session 2, call c (2 GPUs): real two-process ranks over NVLink peer memory with the graph-replayed sharded days; bench at N = 2
O=gpurun_out/c2; mkdir -p $O
nvidia-smi -L
timeout 420 python -m pytest tests/test_gpu_sharded.py -v -rs -x -k "two_ranks" > $O/sharded_2gpu.log 2>&1; echo "sharded rc=$?"
grep -E "PASS|FAIL|ERROR|SKIP|passed|failed|Error" $O/sharded_2gpu.log | tail -12
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open('gpurun_out/c2/bench_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d.get('strong_scaling_256_seeds'), d.get('synth50m'), d.get('single_seed'))
P
tail -3 $O/bench_n2.err
