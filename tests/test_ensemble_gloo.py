"""The N > 1 path on CPU: world_size-2 gloo process group, each rank advancing its own replicas (CPU oracle library
standing in for the GPU engine), final all-reduce of the curve moments == the same ensemble run in one process."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import helpers


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(n_replicas, random_seed):
    return helpers.make_context(helpers.oracle_library(), age_count_override=helpers.small_population(4000),
                                seed=random_seed, n_replicas=n_replicas, max_days=64)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(helpers.ROOT, 'tests'))
    import torch.distributed as dist
    from reina_b200 import ensemble
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    comm = helpers.TorchComm(dist)
    res = ensemble.run_ensemble(_make, days=50, replicas_per_rank=2, seed0=300, comm=comm)
    # percentile bands need the members, not just the moments: all ranks' rows are gathered over the process group
    ctx = res['context']
    rows = ensemble.gather_rows(ctx.series(0, 50), comm)
    bands = ensemble.percentile_bands(rows, (0, 50, 100))
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), mean=res['mean'], std=res['std'], n=res['n'],
             n_rows=rows.shape[0], lo=bands[0], median=bands[50], hi=bands[100])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ensemble_equals_single_process(tmp_path):
    from reina_b200 import ensemble
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / 'rank0.npz'), np.load(tmp_path / 'rank1.npz')
    assert int(r0['n']) == 4 and np.array_equal(r0['mean'], r1['mean']) and np.array_equal(r0['std'], r1['std'])
    ctx = _make(4, 300)                     # seeds 300..303 = rank 0 (300, 301) + rank 1 (302, 303)
    ctx.run(50)
    mean, std = ensemble.mean_std(*ensemble.curve_moments(ctx.series(0, 50)))
    np.testing.assert_allclose(r0['mean'], mean, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(r0['std'], std, rtol=1e-9, atol=1e-6)
    # bands over the gathered members == bands of the one-process ensemble, identical on both ranks
    one = ensemble.percentile_bands(ctx.series(0, 50), (0, 50, 100))
    assert int(r0['n_rows']) == 4
    for name, q in (('lo', 0), ('median', 50), ('hi', 100)):
        assert np.array_equal(r0[name], r1[name])
        np.testing.assert_allclose(r0[name], one[q])
