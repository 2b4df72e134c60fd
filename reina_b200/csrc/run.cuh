// reina_b200 / csrc / run.cuh
// k_run: the whole multi-day run of a FEW replicas as ONE persistent cooperative kernel.
//
// A day is a chain of four dependent phases -- sweep, contacts, resolve, day boundary -- and with few replicas every one
// of them is far too small to fill the GPU: launched as kernels (even from a CUDA graph) a day costs four launch + drain
// + dependent-load latencies, ~57 us for one HUS replica, however little work it holds.  Here every replica gets a TEAM
// of co-resident CTAs for the whole run; the team walks through the phases of day after day with a barrier of its own
// between them (an atomic counter on a private line of the replica's RepCtr, the construction of
// cooperative_groups::grid.sync), so the teams of different replicas never wait for each other.  The phases are the very
// functions the per-phase kernels run (sweep_warp, expose_cta, resolve_part, pre_body / post_body), so the results are
// bit-identical whichever way a run is driven.  The day boundary runs on the team's lead CTA, or -- on days with enough
// capacity events / queued tests (RepCtr::wide_day) -- on a power-of-two sub-team with its own barrier word.
#ifndef REINA_B200_RUN_CUH
#define REINA_B200_RUN_CUH
#include "boundary.cuh"
#include "sweep.cuh"
#include "contacts.cuh"

#define RUN_THREADS 512
#define RUN_WARPS (RUN_THREADS / 32)
#ifndef RUN_CTAS_PER_SM
#define RUN_CTAS_PER_SM 2        // 64 registers
#endif

union RunSmem {
    WarpRings rings[RUN_WARPS];
    struct { int place[8]; uint32_t ri[RUN_WARPS][EX_RCAP], rx[RUN_WARPS][EX_RCAP]; } ex;
    SmemSmall bd;
};

// kind: 0 = start of the first day (k_pre), 1 = end of the last day (k_post), 2 = end of a day + start of the next (k_between).
// ONE call site of post_body / pre_body each (they are thousands of instructions): the team that runs them is picked first.
__device__ __forceinline__ void run_boundary(const Eng &G, const int r, RepCtr *c, RunSmem &S, Team &T, Team &B, const int kind) {
    const bool wide = B.ncta > 1 && (kind == 0 ? c->n_queue >= (uint32_t)G.wide_min : c->wide_day != 0u);
    Team L = team_of(c, 0, 1);
    Team &X = wide ? B : L;
    if (wide ? blockIdx.x < B.ncta : blockIdx.x == 0) {
        if (kind != 0) post_body(G, r, S.bd, X);
        if (kind == 2) team_sync(X);
        if (kind != 1) pre_body(G, r, S.bd, X);
    }
    team_sync(T);
}

__global__ void __launch_bounds__(RUN_THREADS, RUN_CTAS_PER_SM) k_run(Eng G, int n_days, int boundary_ctas) {
    __shared__ RunSmem S;
    const int r = blockIdx.y + G.r0;
    RepCtr *c = &G.ctr[r];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Team T = team_of(c, blockIdx.x, gridDim.x);
    T.bar = &c->run_bar;
    Team B = team_of(c, blockIdx.x, (uint32_t)boundary_ctas);      // barrier word: RepCtr::wide_bar
    // measurement aid (rb_debug_flag 8): nanoseconds the lead CTA spends per phase, barrier included, in RepCtr::dbg_t[8..15]
    const bool timing = G.dbg == 8 && blockIdx.x == 0 && threadIdx.x == 0;
    long long t0 = 0;
    auto lap = [&](int k) { if (timing) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); if (k >= 0) c->dbg_t[8 + k] += t - t0; t0 = t; } };
    lap(-1);
    for (int d = -1; d < n_days; d++) {       // d = -1: only the start of the first day
        const bool last = d == n_days - 1;
        if (d >= 0) {
            sweep_warp(G, r, blockIdx.x * RUN_WARPS + warp, gridDim.x * RUN_WARPS, S.rings[warp], lane);
            lap(1);
            team_sync(T);
            lap(2);
            expose_cta(G, r, blockIdx.x, gridDim.x, S.ex.place, S.ex.ri[warp], S.ex.rx[warp]);
            team_sync(T);
            lap(3);
            if (last) resolve_part<false>(G, r, blockIdx.x * RUN_THREADS + threadIdx.x, gridDim.x * RUN_THREADS);
            else resolve_part<true>(G, r, blockIdx.x * RUN_THREADS + threadIdx.x, gridDim.x * RUN_THREADS);
            team_sync(T);
            lap(4);
        }
        run_boundary(G, r, c, S, T, B, d < 0 ? 0 : (last ? 1 : 2));
        lap(d < 0 ? 0 : 5);
    }
    // Both barrier words back to zero for the next launch -- by the LAST CTA to leave: a CTA counts itself out only after
    // its own wait on the final barrier is over, so when the count is complete nobody polls the words any more.
    if (threadIdx.x == 0 && atomicAdd(&c->run_exit, 1u) == gridDim.x - 1u) { c->run_bar = 0u; c->wide_bar = 0u; c->run_exit = 0u; }
}

#endif
