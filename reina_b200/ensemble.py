"""Monte-Carlo ensembles / scenario sweeps across GPUs (BASELINE configs[3]).

The workload shards as independent units: every rank (one process per GPU) advances its own replicas
(`Context(n_replicas=R, random_seed=seed0 + rank * R)`) with no data-path collective; the only exchange is the
final reduce of the daily curves -- sum and sum of squares per (day, series).  On GPUs that reduce is ONE ncclAllReduce
on device buffers through the engine's own C-ABI (`Context.moments(reduce=True)` -> rb_reduce_moments); percentile
bands gather the members' rows over the same communicator.  Replaces the reference's broken `run_monte_carlo` process
pool (calc/simulation.py:349-385).  No framework is imported here: a communicator is any object with rank / size /
allreduce / allgather / barrier (reina_b200/comm.py; the CPU tests pass an adapter over a gloo group).
"""
import numpy as np

from .comm import LocalComm


def curve_moments(rows):
    """rows[replica, day, series] -> (sum, sum of squares, n) over the replica axis, float64."""
    x = np.asarray(rows, dtype=np.float64)
    return x.sum(axis=0), (x * x).sum(axis=0), x.shape[0]


def reduce_moments(s1, s2, n, comm=None):
    """All-reduce (sum) of host-side curve moments over the communicator, if there is one."""
    comm = comm or LocalComm()
    if comm.size == 1:
        return s1, s2, n
    t = comm.allreduce(np.concatenate([np.ravel(s1), np.ravel(s2), [float(n)]]), 'sum')
    k = s1.size
    return t[:k].reshape(s1.shape), t[k:2 * k].reshape(s2.shape), int(round(t[-1]))


def mean_std(s1, s2, n):
    mean = s1 / n
    var = np.maximum(s2 / n - mean * mean, 0.0) * (n / max(n - 1, 1))
    return mean, np.sqrt(var)


def percentile_bands(rows, q=(5, 25, 50, 75, 95)):
    """Uncertainty bands of an ensemble: rows[replica, day, series] -> {q: [day, series]} (what the reference's UI would
    draw from run_monte_carlo's reina_<scenario>.csv, calc/simulation.py:365-385)."""
    x = np.asarray(rows, dtype=np.float64)
    p = np.percentile(x, list(q), axis=0)
    return {int(k): p[i] for i, k in enumerate(q)}


def gather_rows(rows, comm=None):
    """All ranks' stats rows on every rank (percentiles need the members, not just the moments)."""
    comm = comm or LocalComm()
    rows = np.ascontiguousarray(rows)
    if comm.size == 1:
        return rows
    out = comm.allgather(rows)                       # [rank, replica, day, series]
    return out.reshape((-1,) + rows.shape[1:])


def seeds_for_rank(seed0, replicas_per_rank, rank):
    """Replica r of rank k uses seed seed0 + k * R + r (Context adds r itself)."""
    return seed0 + rank * replicas_per_rank


def run_ensemble(make_context, days, replicas_per_rank, seed0=0, comm=None):
    """Run this rank's share and return the GLOBAL ensemble mean / std of every raw stats column.

    make_context(n_replicas, random_seed) -> reina_b200.model.Context with interventions added.  `comm` None: the
    engine's own NCCL communicator is joined from the launcher's environment (reina_b200.comm.connect) and the reduce
    runs on the device; any other communicator object reduces the host-side moments."""
    from . import comm as _comm
    rank = comm.rank if comm is not None else _comm.world()[0]
    ctx = make_context(replicas_per_rank, seeds_for_rank(seed0, replicas_per_rank, rank))
    if comm is None:
        comm = _comm.connect(ctx._engine)
    ctx.run(days)
    if isinstance(comm, _comm.EngineComm) and comm.engine is ctx._engine:
        s1, s2, n = ctx.moments(0, days, reduce=True)        # device-side: k_moments + one ncclAllReduce
    else:
        s1, s2, n = reduce_moments(*ctx.moments(0, days), comm=comm)
    mean, std = mean_std(s1, s2, n)
    return dict(mean=mean, std=std, n=n, names=ctx.row_layout(), context=ctx, comm=comm)
