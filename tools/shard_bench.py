"""Population-sharded synthetic population (BASELINE configs[4]) on N GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/shard_bench.py [--agents 50e6] [--days 180] [--steps 3] [--warmup 1]
Prints one JSON line (rank 0): whole-job agent-days/s, max over ranks of the device time (CUDA events)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth_run  # noqa: E402
from reina_b200 import inputs, model, sharded  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--agents', type=float, default=50e6)
ap.add_argument('--days', type=int, default=180)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--warmup', type=int, default=1)
ap.add_argument('--force-shard', action='store_true', help='world size 1 through the message / merge path (profiling aid)')
a = ap.parse_args()
rank, world, local = sharded.world()

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))

n = int(a.agents)
f = n / 1685983.0
v = inputs.default_variables()
v['hospital_beds'], v['icu_units'] = int(round(2600 * f)), int(round(300 * f))
ivs = []
for iv in v['interventions']:
    iv = list(iv)
    if iv[0] in ('import-infections', 'import-infections-weekly'):
        iv[2] = int(round(iv[2] * f))
    ivs.append(iv)
v['interventions'] = ivs
args = inputs.build_context_args(v, age_count_override=inputs.synthetic_age_counts(n))
args['random_seed'] = 0
spec = sharded.shard_spec(dist if world > 1 else None) if (world > 1 or a.force_shard) else None
ctx = model.Context(n_replicas=1, device=local, max_days=a.days + 1, shard=spec, **args)
for iv in inputs.active_interventions(v):
    ctx.add_intervention(iv)

ms = []
for step in range(a.warmup + a.steps):
    ctx.reset(1000 + step)              # same seed on every rank
    if world > 1:
        dist.barrier()
    ctx.run(a.days)
    t = torch.tensor([ctx._engine.last_step_ms()], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if step >= a.warmup:
        ms.append(float(t.item()))
rows = ctx.series(0, a.days)[0]
G = len(ctx.age_group_labels)
chk = torch.tensor([float(rows[-1, 3 * G:4 * G].sum())], dtype=torch.float64, device='cuda')
lo, hi = chk.clone(), chk.clone()
if world > 1:
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    m = float(np.mean(ms))
    print(json.dumps(dict(metric='agent-days/sec (synthetic population, population-sharded)', value=n * a.days / (m / 1e3),
                          unit='agent-days/s', n_gpus=world, agents=n, days=a.days, ms_per_run=m, steps=a.steps, warmup=a.warmup,
                          scaling='strong', all_infected_last_day=float(chk.item()),
                          ranks_agree=bool(lo.item() == hi.item()),
                          message_bytes_per_rank_per_day=int(ctx._engine.lib.f['shard_message_bytes'](ctx._engine.h)) if world > 1 else 0,
                          exchange={0: 'none', 1: 'ncclAllGather', 2: 'NVLink peer memory'}[int(ctx._engine.lib.f['shard_exchange'](ctx._engine.h))])),
          flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
