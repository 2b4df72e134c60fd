"""Population-sharded runs (BASELINE configs[4]): one process per GPU, every rank builds the same Context and joins
the others through `shard=(rank, nranks, unique_id)`; the engine exchanges the day's cross-shard events over NVLink peer
memory (include/reina_b200.h, rb_shard_init).

Rank 0's NCCL unique id reaches the other ranks of the node through a file (reina_b200/comm.py); a caller that already
has a process group of its own (`dist`: anything with is_initialized / get_rank / get_world_size / broadcast_object_list)
can hand it over through that instead.
"""
import os

from . import _abi, comm


def world():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by torchrun."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def exchange_unique_id(make_id=None, dist=None):
    """Rank 0 creates the 128-byte NCCL unique id, every rank returns it.  `dist` is an initialised
    torch.distributed module (or None for a single process); `make_id` defaults to the CUDA library's
    rb_shard_unique_id and is injectable so that the plumbing can be tested on CPU."""
    make_id = make_id or _abi.shard_unique_id
    if dist is None or not dist.is_initialized():
        rank, n, _ = world()
        if n == 1:
            return make_id()
        uid = comm.broadcast_bytes(make_id() if rank == 0 else None, rank, key='shard')
        if len(uid) != 128:
            raise _abi.EngineError('unique id hand-off failed')
        return uid
    if dist.get_world_size() == 1:
        return make_id()
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise _abi.EngineError('unique id broadcast failed')
    return bytes(uid)


def shard_spec(dist=None, exchange_capacity=0.0, make_id=None):
    """The `shard=` argument of model.Context for this process."""
    if dist is not None and dist.is_initialized():
        rank, n = dist.get_rank(), dist.get_world_size()
    else:
        rank, n, _ = world()
    return (rank, n, exchange_unique_id(make_id, dist), exchange_capacity)


def join_local(contexts, exchange_capacity=0.0):
    """Several Contexts of THIS process (built from the same inputs, interventions and seed, not yet stepped) become the
    ranks of one population-sharded simulation: rank k = contexts[k].  No NCCL, no other process: the engines read each
    other's message buffers directly -- one device (each rank on its own stream; how a single-GPU box exercises the
    multi-rank exchange) or several devices with peer access."""
    _abi.shard_init_local([c._engine for c in contexts], exchange_capacity)
    return contexts


def run_local(contexts, days):
    """Advance every local rank by `days`.  A rank's day ends only once every rank's sweep of that day has been launched,
    and a rank that is waiting occupies the device: so everything that may allocate, upload or synchronise (planning,
    contact tables, the schedule) is done for all ranks first, and only the asynchronous launches run side by side, one
    host thread per rank."""
    import threading
    for c in contexts:
        c._prepare_run(days)
    for c in contexts:
        c._engine.sync()
    errors = []

    def work(ctx):
        try:
            ctx._launch_run(days)
            ctx._finish_run()
        except Exception as exc:      # noqa: BLE001 -- re-raised below, in the caller's thread
            errors.append(exc)
    threads = [threading.Thread(target=work, args=(c,)) for c in contexts]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]


def merge_agents(per_rank_agents):
    """Canonical agent records of the whole population from every rank's rb_read_agents: agent a is taken from the
    rank that owns it (day counters and severity are authoritative there)."""
    import numpy as np
    n = len(per_rank_agents)
    out = per_rank_agents[0].copy()
    owner = _abi.owner_of(np.arange(len(out)), n)
    for r in range(1, n):
        sel = owner == r
        out[sel] = per_rank_agents[r][sel]
    return out
