#!/bin/bash
# round 2, session 2, call h (8 GPUs): the default bench line at N = 8 with the final code (weak scaling, the 256-seed strong-scaling
# point with 4 half-wave groups + launch priorities, the synthetic 50 M population sharded over 8 GPUs with graph-replayed days)
O=gpurun_out/n8; mkdir -p $O
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open('gpurun_out/n8/bench_n8.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d.get('strong_scaling_256_seeds'), d.get('synth50m'))
P
tail -2 $O/bench_n8.err
