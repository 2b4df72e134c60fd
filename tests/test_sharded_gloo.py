"""Host side of the population-sharded mode on CPU: world_size-2 gloo group hands rank 0's unique id to every rank
(reina_b200.sharded), and per-rank agent reads are merged by ownership.  The device side (NCCL all-gather, k_merge) is
covered by tests/test_gpu_sharded.py on the GPU box."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import helpers
from reina_b200 import _abi, sharded


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(helpers.ROOT, 'tests'))
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # a stand-in id factory: ncclGetUniqueId needs no GPU, but the point here is the plumbing
    make_id = lambda: bytes([(7 * i + 3 * rank + 1) % 256 for i in range(128)])
    spec = sharded.shard_spec(dist, exchange_capacity=2.0, make_id=make_id)
    with open(os.path.join(out_dir, 'rank%d.bin' % rank), 'wb') as f:
        f.write(bytes([spec[0], spec[1]]) + spec[2])
    assert spec[3] == 2.0
    dist.barrier()
    dist.destroy_process_group()


def test_unique_id_reaches_every_rank(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    b0, b1 = (open(tmp_path / ('rank%d.bin' % r), 'rb').read() for r in range(2))
    assert b0[0] == 0 and b1[0] == 1 and b0[1] == b1[1] == 2
    assert b0[2:] == b1[2:] == bytes([(7 * i + 1) % 256 for i in range(128)])      # rank 0's id, on both ranks


def test_single_process_spec_and_agent_merge():
    spec = sharded.shard_spec(None, make_id=lambda: b'\x05' * 128)
    assert spec[:2] == (0, 1) and spec[2] == b'\x05' * 128
    n = 3 * 4096 + 100
    per_rank = []
    for r in range(3):
        a = np.zeros(n, dtype=_abi.AGENT_DTYPE)
        a['days_left'] = r + 1
        per_rank.append(a)
    merged = sharded.merge_agents(per_rank)
    owner = _abi.owner_of(np.arange(n), 3)
    assert np.array_equal(merged['days_left'], owner + 1)
    assert list(np.unique(owner[:4096])) == [0] and owner[4096] == 1 and owner[2 * 4096] == 2 and owner[3 * 4096] == 0
