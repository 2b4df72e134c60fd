#!/usr/bin/env python
"""bench.py -- agent-days/sec of the per-day agent loop on the HUS configuration (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--replicas R] [--days D] [--impl reference]

A "step" is one complete HUS run (1,685,983 agents x 180 days, the reference's default interventions) of an
ensemble of R seeds per GPU, every step with fresh seeds.  `value` times only the device-resident multi-day run
(CUDA events on the engine's stream, inputs already in HBM); `e2e` times the same step through the public API
(`Context.reset` + `upload_inputs` + `run` + `series`) with host buffers, host->device and device->host copies
included.  For N > 1 (torchrun, one rank per GPU) the ensemble is partitioned over the GPUs -- independent units,
weak scaling -- and the only collective is the final NCCL all-reduce of the daily curves (sum and sum of squares).

`--impl reference` times the UNMODIFIED reference engine (oracle/_ref, built from /root/reference by
oracle/build_ref.sh) on this box's host cores: one seed per core, same workload and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AREA = 'HUS'
N_AGENTS = 1685983


def load_sweep_traffic(n_replicas):
    """Mean DRAM bytes (read + write) per k_sweep launch over the 180 launches of one run, from the committed ncu pass
    (profiles/: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` on every k_sweep launch,
    tools/collect_profiles.sh).  Only valid for the replica count it was captured at."""
    p = os.path.join(ROOT, 'profiles', 'r01_launches_and_sweep_traffic_R%d.json' % n_replicas)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return float(json.load(f)['k_sweep_dram']['mean_traffic_bytes'])


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    samples=len(sm), reasons=sorted(reasons))


def make_context(n_replicas, device, days, seed):
    from reina_b200 import inputs, model
    v = inputs.default_variables()
    args = inputs.build_context_args(v, area=AREA)
    args['random_seed'] = seed
    ctx = model.Context(n_replicas=n_replicas, device=device, max_days=days + 1, **args)
    for iv in inputs.active_interventions(v):
        ctx.add_intervention(iv)
    return ctx


def algorithmic_bytes(s1, n_agents, n_replicas, G):
    """bytes_day = 4 N + 12 I_d + 8 E_d, summed over days and replicas.
    s1: [D, row_len] sums over replicas of the stats rows (I_d = infected that day, E_d = contacts sampled that day)."""
    from reina_b200 import _abi
    nA = len(_abi.ATTRS)
    i_inf = _abi.ATTRS.index('infected')
    I = s1[:, i_inf * G:(i_inf + 1) * G].sum(axis=1)                     # state at start of day d, all replicas
    E = s1[:, nA * G + _abi.SCALARS.index('exposed_per_day')]            # contacts of day d-1, all replicas
    D = s1.shape[0]
    sweep = 4.0 * n_agents * n_replicas * D + 12.0 * I.sum()
    total = sweep + 8.0 * E.sum()
    return dict(sweep=sweep, total=total, mean_infected=float(I.mean() / n_replicas),
                mean_contacts=float(E[1:].mean() / n_replicas) if D > 1 else 0.0,
                daily_infected=(I / n_replicas).tolist(), daily_contacts=(E / n_replicas).tolist())


def cpu_baseline(days, seeds, processes):
    """The reference's own Cython engine (oracle/_ref) -- or, if it is not built, the C oracle port -- timed on
    this box's host cores: agent-days/s over the summed iterate() time of a bounded sample."""
    from oracle import ref_harness
    if ref_harness.available():
        t0 = time.perf_counter()
        _, t_iter, wall = ref_harness.run_ensemble(seeds, days=days, processes=processes, area=AREA)
        # rate of each core = agent-days of its run / time spent inside iterate() (setup excluded); cores add up
        value = float((N_AGENTS * days / t_iter).sum()) if processes > 1 else N_AGENTS * days * len(seeds) / float(t_iter.sum())
        return dict(value=value, unit='agent-days/s', cores=processes, kind='reference',
                    sample='%d seed(s) x HUS %d days, unmodified cythonsim engine (oracle/_ref), %s'
                           % (len(seeds), days, 'sum of iterate() time' if processes == 1 else 'sum over cores of agent-days / iterate() time'),
                    seconds=time.perf_counter() - t0)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import helpers
    t0 = time.perf_counter()
    ctx = helpers.make_context(helpers.oracle_library(), area=AREA, seed=int(seeds[0]), max_days=days + 1)
    ctx.run(days)
    t = time.perf_counter() - t0
    return dict(value=N_AGENTS * days / t, unit='agent-days/s', cores=1, kind='port',
                sample='1 seed x HUS %d days, C oracle port (oracle/_ref not built)' % days, seconds=t)


def run_reference(a, rank, world):
    if rank != 0:
        return
    procs = min(os.cpu_count() or 1, 32)
    vals = []
    for step in range(a.warmup + a.steps):
        if step < a.warmup and step > 0:
            continue        # one warm-up pass is enough to page the extension in; each pass costs ~15 s
        seeds = np.arange(procs) + 100000 + 1000 * step
        b = cpu_baseline(a.days, seeds, procs)
        if step >= a.warmup:
            vals.append(b)
    value = float(np.mean([b['value'] for b in vals]))
    ms = 1e3 * N_AGENTS * a.days * procs / value
    out = dict(metric='agent-days/sec (HUS 1.7M)', value=value, unit='agent-days/s', impl='reference',
               n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=ms, higher_is_better=True,
               scaling='weak', vs_baseline=None, dtype='int32/f32 state, f64 uniforms', data='synthetic',
               config=dict(workload='HUS 1,685,983 agents x %d days, default interventions, %d seeds per step (one per host core)'
                                    % (a.days, procs)),
               cpu_baseline=dict(value=value, unit='agent-days/s', cores=procs, kind=vals[0]['kind'], sample=vals[0]['sample']),
               e2e=dict(value=value, unit='agent-days/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--replicas', type=int, default=256, help='ensemble members (seeds) per GPU (BASELINE configs[3]: 256 seeds)')
    ap.add_argument('--days', type=int, default=180)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-single', action='store_true')
    ap.add_argument('--dump-daily', default=None, help='write the per-day I_d, E_d (ensemble means) the algorithmic bytes are computed from')
    a = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if a.impl == 'reference':
        run_reference(a, rank, world)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    R, D = a.replicas, a.days
    peak_gbs, peak_src = load_peaks()
    ctx = make_context(R, local, D, seed=1)
    eng = ctx._engine
    G = len(ctx.age_group_labels)
    seed_of = lambda step: 1_000_000 * (rank + 1) + 1000 * step     # fresh seeds every step, distinct per rank

    from reina_b200 import ensemble

    # ---------------- device-resident arm (`value`) and end-to-end arm (`e2e`), same steps ----------------
    def one_step(step, timed):
        t0 = time.perf_counter()
        ctx.reset(seed_of(step))                   # fresh ensemble; device state re-initialised in place
        h2d = ctx.upload_inputs()                  # contact tables of every mobility epoch, from host memory
        ctx.run(D)                                 # schedule H2D + 180 simulated days + sync
        dev_ms = eng.last_step_ms()                # CUDA events around the 180-day run only
        s1, s2, n = ctx.moments(0, D)              # the step's result: sum / sum of squares of every daily series
        s1g, s2g, ng = ensemble.reduce_moments(s1, s2, n)      # final reduce over the GPUs (NCCL all-reduce)
        wall = time.perf_counter() - t0
        h2d += D * 256                             # sizeof(rb_day_params) per day
        return dev_ms, wall, s1, h2d, s1.nbytes + s2.nbytes

    for step in range(a.warmup):
        one_step(step, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    barrier()
    dev_ms_total, wall_total = 0.0, 0.0
    s1 = None
    for step in range(a.warmup, a.warmup + a.steps):
        dev_ms, wall, s1, h2d, d2h = one_step(step, True)
        dev_ms_total += dev_ms
        wall_total += wall
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    dev_ms_total = max_over_ranks(dev_ms_total)
    wall_total = max_over_ranks(wall_total)

    agent_days_step = float(N_AGENTS) * D * R * world
    ms_per_step = dev_ms_total / a.steps
    value = agent_days_step / (ms_per_step / 1e3)
    e2e_value = agent_days_step / (wall_total / a.steps)
    alg = algorithmic_bytes(s1, N_AGENTS, R, G)

    if a.dump_daily and rank == 0:
        with open(a.dump_daily, 'w') as f:
            json.dump(dict(note='HUS, ensemble means over %d seeds of the last timed step: I_d = infected at the start of day d, '
                                'E_d = contacts sampled on day d-1 (stats row d); bytes_d = 4 N + 12 I_d + 8 E_d (SURVEY.md 8d)' % R,
                           agents=N_AGENTS, days=D, I_d=alg['daily_infected'], E_d=alg['daily_contacts']), f)

    # ---------------- per-kernel device times (one extra run, events around every launch) ----------------
    ctx.reset(seed_of(a.warmup + a.steps - 1))
    while len(ctx._plan) < D:
        ctx._plan_next_day()
    eng.set_schedule(0, ctx._plan[:D])
    kms = eng.step_profiled(D)
    knames = ['k_pre', 'k_sweep', 'k_expose', 'k_resolve', 'k_post']
    ksum = float(kms.sum())
    sweep_ms = float(kms[1]) / D
    roof_achieved = alg['sweep'] / D / (sweep_ms / 1e3) / 1e9           # GB/s, algorithmic bytes of one sweep launch
    whole = alg['total'] / (ms_per_step / 1e3) / 1e9

    out = dict(
        metric='agent-days/sec (HUS 1.7M)', value=value, unit='agent-days/s', n_gpus=world, steps=a.steps,
        warmup=a.warmup, ms_per_step=ms_per_step, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='int32/f32 state, f64 uniforms', data='synthetic',
        config=dict(
            workload='HUS 1,685,983 agents x %d days, default interventions (BASELINE configs[1]), ensemble of %d seeds per GPU '
                     'advanced by the same launches (configs[3] share)' % (D, R),
            replicas_per_gpu=R, days=D, agents=N_AGENTS, parallelism='ensemble x%d' % world,
            l2='inputs larger than L2 (%.1f GB of agent state per GPU)' % (R * N_AGENTS * 36.3 / 1e9) if R > 2 else
               'single 6.7 MB packed-state array is L2-resident by nature of the workload (180 dependent days)',
        ),
        e2e=dict(value=e2e_value, unit='agent-days/s', h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                 ms_per_step=1e3 * wall_total / a.steps),
        gpu_launches=int(launches),
        clocks=clocks,
        roofline=dict(bound='hbm', kernel='k_sweep', achieved=roof_achieved, peak=peak_gbs, unit='GB/s',
                      frac=roof_achieved / peak_gbs, traffic=load_sweep_traffic(R) if D == 180 else None, peak_source=peak_src,
                      algorithmic_bytes_per_launch=alg['sweep'] / D, avg_launch_ms=sweep_ms,
                      share_of_step=float(kms[1]) / ksum),
        roofline_whole_run=dict(achieved=whole, peak=peak_gbs, unit='GB/s', frac=whole / peak_gbs,
                                bytes_per_agent_day=alg['total'] / (float(N_AGENTS) * D * R),
                                mean_infected=alg['mean_infected'], mean_contacts_per_day=alg['mean_contacts']),
        kernel_ms_per_day={k: float(v) / D for k, v in zip(knames, kms)},
    )

    # ---------------- single-seed run of the same configuration (latency-bound, reported beside) ----------------
    if not a.no_single and rank == 0 and R != 1:
        ctx.close()
        c1 = make_context(1, local, D, seed=1)
        for s in range(2):
            c1.reset(50 + s); c1.run(D)
        ms = []
        for s in range(3):
            c1.reset(60 + s); c1.run(D); ms.append(c1._engine.last_step_ms())
        out['single_seed'] = dict(value=N_AGENTS * D / (np.mean(ms) / 1e3), unit='agent-days/s',
                                  ms_per_run=float(np.mean(ms)), us_per_day=1e3 * float(np.mean(ms)) / D)
        c1.close()

    if rank == 0 and not a.no_cpu_baseline and world == 1:
        b = cpu_baseline(D, np.array([0]), 1)
        out['cpu_baseline'] = {k: b[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    elif rank == 0:
        out['cpu_baseline'] = None

    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
