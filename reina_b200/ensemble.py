"""Monte-Carlo ensembles / scenario sweeps across GPUs (BASELINE configs[3]).

The workload shards as independent units: every rank (one process per GPU) advances its own replicas
(`Context(n_replicas=R, random_seed=seed0 + rank * R)`) with no data-path collective; the only exchange is the
final reduce of the daily curves -- sum and sum of squares per (day, series) -- over `torch.distributed`
(NCCL on GPUs, gloo in the CPU tests).  Replaces the reference's broken `run_monte_carlo` process pool
(calc/simulation.py:349-385).  torch is plumbing only (process group + all_reduce); without an initialised
process group everything here is plain numpy.
"""
import numpy as np


def curve_moments(rows):
    """rows[replica, day, series] -> (sum, sum of squares, n) over the replica axis, float64."""
    x = np.asarray(rows, dtype=np.float64)
    return x.sum(axis=0), (x * x).sum(axis=0), x.shape[0]


def _dist():
    try:
        import torch.distributed as dist
    except Exception:
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def reduce_moments(s1, s2, n):
    """All-reduce (sum) of the curve moments over the process group, if there is one."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return s1, s2, n
    import torch
    t = torch.from_numpy(np.concatenate([s1.ravel(), s2.ravel(), [float(n)]]))
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    dist.all_reduce(t)
    t = t.cpu().numpy()
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


def gather_rows(rows):
    """All ranks' stats rows on every rank (percentiles need the members, not just the moments): all_gather over the
    process group, identity without one."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return np.asarray(rows)
    import torch
    t = torch.from_numpy(np.ascontiguousarray(rows))
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy() for o in out], axis=0)


def seeds_for_rank(seed0, replicas_per_rank, rank):
    """Replica r of rank k uses seed seed0 + k * R + r (Context adds r itself)."""
    return seed0 + rank * replicas_per_rank


def run_ensemble(make_context, days, replicas_per_rank, seed0=0, rank=0):
    """Run this rank's share and return the GLOBAL ensemble mean / std of every raw stats column.

    make_context(n_replicas, random_seed) -> reina_b200.model.Context with interventions added."""
    ctx = make_context(replicas_per_rank, seeds_for_rank(seed0, replicas_per_rank, rank))
    ctx.run(days)
    s1, s2, n = reduce_moments(*ctx.moments(0, days))      # reduced over replicas on the device, over ranks by NCCL/gloo
    mean, std = mean_std(s1, s2, n)
    return dict(mean=mean, std=std, n=n, names=ctx.row_layout())
