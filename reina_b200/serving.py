"""Serving hook: a long-lived simulation worker over resident GPU contexts.

The reference forks one `multiprocessing.Process` per request (simulation_thread.py:14-61): the child builds a fresh
`model.Context` (allocating and initialising 1.7 M agents), runs `simulate_individuals` with a `step_callback` that
publishes partial results under `<cache_key>-results`, and flags `<cache_key>-finished` / `<cache_key>-error`;
`graphql_schema.RunSimulation` (:382-408) returns the process uuid and the UI polls the cache.  A GPU run takes
milliseconds, so forking per request would cost far more than the run: here ONE worker thread owns the contexts, keeps
them allocated between requests (`Context.reset` re-initialises the device state in place; contexts are keyed by
everything that shapes the allocation or the schedule), and serves requests from a queue.

    worker = SimulationWorker(device=0)             # SimulationProcess(variables).start() becomes:
    job = worker.submit(variables)                  #   -> job id (the reference's process uuid)
    res = worker.results(job)                       #   -> dict(total=df, age_groups=adf, finished=bool, error=str | None)
    worker.wait(job); worker.close()

`results()` carries the same three cache entries the reference's poller reads (results / finished / error).  Partial
results appear every `callback_day_interval` simulated days, as with the reference's step_callback.
"""
import json
import queue
import threading
import uuid

from . import inputs, model, simulation


def _context_key(variables, scenario):
    """Everything that shapes the device allocation or the host-side schedule of a Context."""
    v = {k: variables[k] for k in sorted(variables) if k != 'random_seed'}
    return json.dumps([v, scenario], sort_keys=True, default=str)


class SimulationWorker:
    def __init__(self, device=0, max_contexts=4, callback_day_interval=30, context_factory=None, max_finished_jobs=256):
        self.device = device
        self.max_contexts = max_contexts
        # finished jobs are kept for polling clients, the oldest are dropped beyond this many (the reference's cache
        # entries expire, simulation_thread.py:40-60); release(job) drops one at once
        self.max_finished_jobs = max_finished_jobs
        self._finished = []
        self.callback_day_interval = callback_day_interval
        # injectable so that the host logic can be tested without a GPU (tests pass the CPU oracle's library)
        self._factory = context_factory or (lambda v, scenario: simulation.make_context(v, device=self.device, scenario=scenario))
        self._contexts = {}          # key -> Context, in least-recently-used order
        self._jobs = {}
        self._lock = threading.Lock()
        self._queue = queue.Queue()
        self._thread = threading.Thread(target=self._serve, daemon=True)
        self._thread.start()

    # -- client side ------------------------------------------------------------------------------
    def submit(self, variables=None, scenario=None):
        v = dict(variables or inputs.default_variables())
        job = str(uuid.uuid4())
        with self._lock:
            self._jobs[job] = dict(total=None, age_groups=None, finished=False, error=None, done=threading.Event(),
                                   cancelled=False, reused_context=None)
        self._queue.put((job, v, scenario))
        return job

    def results(self, job):
        with self._lock:
            j = self._jobs[job]
            return dict(total=j['total'], age_groups=j['age_groups'], finished=j['finished'], error=j['error'],
                        reused_context=j['reused_context'])

    def cancel(self, job):
        """The reference's step_callback returning False (ExecutionInterrupted, calc/simulation.py:282-284)."""
        with self._lock:
            self._jobs[job]['cancelled'] = True

    def wait(self, job, timeout=None):
        with self._lock:
            done = self._jobs[job]['done']
        done.wait(timeout)
        return self.results(job)

    def release(self, job):
        """Forget a job and its DataFrames (a finished one, or one whose results nobody will ask for)."""
        with self._lock:
            self._jobs.pop(job, None)
            if job in self._finished:
                self._finished.remove(job)

    def close(self):
        self._queue.put(None)
        self._thread.join()
        for ctx in self._contexts.values():
            ctx.close()
        self._contexts.clear()

    # -- worker thread ----------------------------------------------------------------------------
    def _context_for(self, v, scenario):
        key = _context_key(v, scenario)
        ctx = self._contexts.pop(key, None)
        reused = ctx is not None
        if ctx is None:
            while len(self._contexts) >= self.max_contexts:
                self._contexts.pop(next(iter(self._contexts))).close()       # evict the least recently used
            ctx = self._factory(v, scenario)
        else:
            ctx.reset(v['random_seed'])
        self._contexts[key] = ctx
        return ctx, reused

    def _serve(self):
        while True:
            item = self._queue.get()
            if item is None:
                return
            job, v, scenario = item
            with self._lock:
                j = self._jobs.get(job)
            if j is None:                   # released before it started
                continue
            try:
                ctx, reused = self._context_for(v, scenario)
                j['reused_context'] = reused

                def step_callback(part, _j=j):
                    with self._lock:
                        _j['total'] = part
                        return not _j['cancelled']

                df, adf = simulation.simulate_individuals(v, step_callback=step_callback,
                                                          callback_day_interval=self.callback_day_interval, context=ctx)
                with self._lock:
                    j['total'], j['age_groups'] = df, adf
            except simulation.ExecutionInterrupted:
                with self._lock:
                    j['error'] = 'cancelled'
            except Exception as e:      # the reference stores str(e) under <key>-error (simulation_thread.py:52-55)
                with self._lock:
                    j['error'] = str(e) or type(e).__name__
            with self._lock:
                j['finished'] = True
                if job in self._jobs:
                    self._finished.append(job)
                while len(self._finished) > self.max_finished_jobs:
                    self._jobs.pop(self._finished.pop(0), None)
            j['done'].set()
