"""Population-sharded mode (BASELINE configs[4]): N ranks, one process per GPU, one NCCL all-gather per day.

Every rank must reproduce the sequential CPU oracle bit for bit: identical daily series on every rank, identical test
queue and capacity counters, and every agent (taken from the rank that owns it) identical to the oracle's.
world_size 1 exercises the message / merge path on the single-GPU box; world_size 2+ needs that many GPUs
(run under `gpurun --gpus N`)."""
import multiprocessing as mp
import os
import traceback

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

# fields of rb_agent every rank keeps exact for every agent (the rest is authoritative on the owner only)
REPLICATED_FIELDS = ('infector', 'n_infected', 'day_of_vaccination', 'state', 'variant')


def _case(name):
    v = helpers.inputs.default_variables()
    if name == 'stress':
        counts = helpers.small_population(80000)
        v['hospital_beds'], v['icu_units'] = 25, 3
        return dict(variables=v, age_count_override=counts, seed=11, interventions=helpers.stress_interventions()), 120
    if name == 'default':
        counts = helpers.small_population(150000)
        v['hospital_beds'], v['icu_units'] = helpers.scaled_capacity(150000)
        return dict(variables=v, age_count_override=counts, seed=2), 180
    if name == 'tracing':
        counts = helpers.small_population(120000)
        v['hospital_beds'], v['icu_units'] = helpers.scaled_capacity(120000)
        return dict(variables=v, age_count_override=counts, seed=5, scenario='hammer-and-dance'), 180
    raise KeyError(name)


def _worker(rank, world, uid, case, chunk, q):
    try:
        kw, days = _case(case)
        ctx = helpers.make_context(helpers.cuda_library(), device=rank, shard=(rank, world, uid), **kw)
        done = 0
        while done < days:
            n = min(chunk, days - done)
            ctx.run(n)
            done += n
        eng = ctx._engine
        q.put((rank, dict(series=ctx.series(0, days), agents=eng.read_agents(0), queue=np.sort(eng.read_queue(0)),
                          avail=eng.read_available(0), launches=eng.launch_count(),
                          msg_bytes=eng.lib.f['shard_message_bytes'](eng.h), exchange=eng.lib.f['shard_exchange'](eng.h))))
    except Exception:
        q.put((rank, traceback.format_exc()))


def _run_sharded(world, case, chunk=1000):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    uid = helpers._abi.shard_unique_id()
    mpc = mp.get_context('spawn')
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, uid, case, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in range(world):
            rank, res = q.get(timeout=600)
            assert not isinstance(res, str), 'rank %d failed:\n%s' % (rank, res)
            out[rank] = res
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    return out


def _check_against_oracle(out, world, case):
    kw, days = _case(case)
    cpu = helpers.make_context(helpers.oracle_library(), **kw)
    cpu.run(days)
    ref = cpu.series(0, days)
    names = cpu.row_layout()
    ca = cpu._engine.read_agents(0)
    owner = helpers._abi.owner_of(np.arange(len(ca)), world)
    for rank in range(world):
        res = out[rank]
        bad = np.argwhere(res['series'] != ref)
        assert len(bad) == 0, 'rank %d: series differ at %d cells, first day %d column %s: %d vs oracle %d' % (
            rank, len(bad), bad[0][1], names[bad[0][2]], res['series'][tuple(bad[0])], ref[tuple(bad[0])])
        assert np.array_equal(res['queue'], np.sort(cpu._engine.read_queue(0))), 'rank %d: test queue differs' % rank
        assert np.array_equal(res['avail'], cpu._engine.read_available(0)), 'rank %d: free beds / ICU differ' % rank
        mine = owner == rank
        for f in ca.dtype.names:
            sel = slice(None) if f in REPLICATED_FIELDS else mine
            a, b = res['agents'][f][sel], ca[f][sel]
            assert np.array_equal(a, b), 'rank %d: agent field %s differs for %d agents' % (rank, f, int((a != b).sum()))
        # detected / queued / has-list flags are replicated too (bit 2, included_in_totals, is the owner's)
        assert np.array_equal(res['agents']['flags'] & 0b1011, ca['flags'] & 0b1011), 'rank %d: replicated flags differ' % rank
    G = len(cpu.age_group_labels)
    assert ref[0, -1, 3 * G:4 * G].sum() > 0.02 * cpu.n_agents      # the epidemic happened


@pytest.mark.parametrize('case', ['stress', 'default'])
def test_world_size_1_message_path(case):
    """One rank: the sweep writes its message, the all-gather is the identity, k_merge applies it."""
    out = _run_sharded(1, case, chunk=50)
    _check_against_oracle(out, 1, case)


@pytest.fixture
def wide_boundary(monkeypatch):
    """Force the multi-CTA day boundary on every day (the spawned rank processes inherit the environment)."""
    monkeypatch.setenv('RB_WIDE_MIN', '0')
    monkeypatch.setenv('RB_WIDE_CTAS', '10')


@pytest.mark.parametrize('case', ['stress', 'tracing'])
def test_world_size_1_wide_boundary(wide_boundary, case):
    out = _run_sharded(1, case, chunk=50)
    _check_against_oracle(out, 1, case)


def test_two_ranks_wide_boundary(wide_boundary):
    out = _run_sharded(2, 'stress', chunk=31)
    _check_against_oracle(out, 2, 'stress')


@pytest.mark.parametrize('case', ['stress', 'default', 'tracing'])
def test_two_ranks_equal_oracle(case):
    out = _run_sharded(2, case, chunk=31)
    _check_against_oracle(out, 2, case)
    assert out[0]['msg_bytes'] > 0
    assert out[0]['exchange'] == out[1]['exchange'] == 2, 'expected the NVLink peer-memory exchange between two processes'


def test_two_ranks_nccl_exchange(monkeypatch):
    """The fallback exchange: fixed-size slots through one ncclAllGather per day."""
    monkeypatch.setenv('RB_SHARD_EXCHANGE', 'nccl')
    out = _run_sharded(2, 'stress', chunk=31)
    _check_against_oracle(out, 2, 'stress')
    assert out[0]['exchange'] == 1


def test_four_ranks_equal_oracle():
    out = _run_sharded(4, 'stress')
    _check_against_oracle(out, 4, 'stress')


def test_eight_ranks_equal_oracle():
    out = _run_sharded(8, 'default')
    _check_against_oracle(out, 8, 'default')


# ---------------------------------------------------------------- several ranks on ONE GPU (rb_shard_init_local)
# The ranks of one process, each engine on its own stream of the same device, read each other's message buffers through
# plain device pointers: the flags, k_wait, k_merge over several messages and the ownership split run exactly as they do
# across GPUs, so a single-GPU box checks the multi-rank exchange too.  A spawned process, so that
# CUDA_DEVICE_MAX_CONNECTIONS (one hardware queue per rank: a waiting rank must never sit in front of a peer's kernels)
# is in force when the CUDA context is created.
def _local_worker(world, case, chunk, q):
    try:
        from reina_b200 import sharded
        kw, days = _case(case)
        ctxs = [helpers.make_context(helpers.cuda_library(), device=0, **kw) for _ in range(world)]
        sharded.join_local(ctxs)
        done = 0
        while done < days:
            n = min(chunk, days - done)
            sharded.run_local(ctxs, n)
            done += n
        out = {}
        for rank, ctx in enumerate(ctxs):
            eng = ctx._engine
            out[rank] = dict(series=ctx.series(0, days), agents=eng.read_agents(0), queue=np.sort(eng.read_queue(0)),
                             avail=eng.read_available(0), launches=eng.launch_count(),
                             msg_bytes=eng.lib.f['shard_message_bytes'](eng.h), exchange=eng.lib.f['shard_exchange'](eng.h))
        q.put(out)
    except Exception:
        q.put(traceback.format_exc())


def _run_local_ranks(world, case, chunk, monkeypatch):
    monkeypatch.setenv('CUDA_DEVICE_MAX_CONNECTIONS', '32')
    monkeypatch.setenv('CUDA_MODULE_LOADING', 'EAGER')      # belt and braces: rb_shard_init_local loads its kernels itself
    mpc = mp.get_context('spawn')
    q = mpc.Queue()
    p = mpc.Process(target=_local_worker, args=(world, case, chunk, q))
    p.start()
    try:
        out = q.get(timeout=300)
    finally:
        p.join(timeout=30)
        if p.is_alive():
            p.kill()
    assert not isinstance(out, str), 'local ranks failed:\n%s' % out
    return out


@pytest.mark.parametrize('world,case,chunk', [(2, 'stress', 31), (3, 'tracing', 1000), (4, 'default', 50), (8, 'stress', 1000)])
def test_local_ranks_on_one_gpu_equal_oracle(monkeypatch, world, case, chunk):
    out = _run_local_ranks(world, case, chunk, monkeypatch)
    _check_against_oracle(out, world, case)
    assert all(out[r]['exchange'] == 2 for r in range(world))
