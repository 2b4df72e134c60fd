// reina_b200 / csrc / shard.cuh
// k_merge: population-sharded mode, applies every rank's message of the day (see include/reina_b200.h, rb_shard_init).
#ifndef REINA_B200_SHARD_CUH
#define REINA_B200_SHARD_CUH
#include "state.cuh"

// ---------------------------------------------------------------- peer exchange: publish / wait
// Exchange through peer memory (Eng::xp2p): once its sweep and contact kernels are done, a rank raises the flag at the
// head of its own buffer to xflag_value(day); every rank waits (k_wait) for every owner's flag, and k_merge then reads the
// message straight out of the owner's memory over NVLink -- header first, then exactly as many list entries as the
// header counts -- instead of receiving fixed-size slots from an all-gather.  The slots alternate with the day's
// parity: a rank overwrites the slot of day d on day d+2, after its day-d+1 merge, which waited for every peer's
// day-d+1 flag, which every peer raised after finishing its day-d merge.
__global__ void k_publish(Eng G) {
    if (threadIdx.x != 0) return;
    __threadfence_system();                                   // the kernels before this one wrote the message
    *(volatile uint32_t *)G.xpeer[G.rank] = xflag_value(G, G.ctr[0].day);
}
// One warp, lane k watches rank k's flag: a single poller per peer keeps the NVLink request queues free for the data.
__global__ void k_wait(Eng G) {
    RepCtr *c = &G.ctr[0];
    if (c->problem) return;                                   // the run has failed already (sticky): do not wait a minute per remaining day
    if ((int)threadIdx.x < G.nranks) {
        const volatile uint32_t *flag = (const volatile uint32_t *)G.xpeer[threadIdx.x];
        const uint32_t want = xflag_value(G, c->day);
        const long long t0 = clock64();
        while ((int32_t)(*flag - want) < 0) {
            __nanosleep(200);
            if (clock64() - t0 > 120000000000ll) { set_problem(c, RB_OTHER_FAILURE); break; }    // ~1 min: a peer died; fail loudly, do not hang
        }
        __threadfence_system();
    }
}
// message data is read once, from the owner's L2 (never through this SM's L1)
__device__ __forceinline__ uint32_t pull(const uint32_t *p) { return __ldcg(p); }
__device__ __forceinline__ int32_t pull(const int32_t *p) { return __ldcg(p); }
__device__ __forceinline__ unsigned long long pull(const unsigned long long *p) { return __ldcg(p); }
__device__ __forceinline__ uint2 pull(const uint2 *p) { return __ldcg(p); }
__device__ __forceinline__ Attempt pull(const Attempt *p) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p));
    Attempt a; a.cand = v.x; a.parent = v.y; a.key = ((unsigned long long)v.w << 32) | v.z;
    return a;
}

// ---------------------------------------------------------------- k_merge (population-sharded mode only)
// After the exchange every rank can see every rank's message.  All ranks apply all of them in rank order, so the
// replicated state (counters, test queue, capacity events, packed words, conflict slots) stays identical everywhere:
// count deltas are added, queue entries / events / successful transmissions are concatenated into the single-GPU
// lists, the other ranks' state changes overwrite the local copies of their agents, and every successful
// transmission does its atomicMin on the target's conflict slot (first infector in sweep order wins, main.pyx:238-244).
// Reading a peer's memory costs an NVLink round trip (microseconds), so the kernel is laid out to need few of them in
// sequence: the list lengths of all ranks are fetched by one thread each, the blocks are dealt to the ranks and walk the
// four lists of "their" rank as ONE index space (independent load -> store pairs), and the counter deltas are summed by
// one thread per counter across the whole grid.  gridDim.x is a multiple of nranks.
__global__ void __launch_bounds__(256) k_merge(Eng G) {
    __shared__ uint32_t raw[MAX_RANKS][4], cnt[MAX_RANKS][4], off[MAX_RANKS + 1][4];     // n_newq, n_events, n_upd, n_succ per rank
    RepCtr *c = &G.ctr[0];
    const int nrk = G.nranks;
    const int day = c->day;
    if ((int)threadIdx.x < nrk * 4) {
        const int k = threadIdx.x >> 2, f = threadIdx.x & 3;
        const RepCtr *h = xslot_of(G, k, day).hdr;
        const uint32_t *p = f == 0 ? &h->n_newq : (f == 1 ? &h->n_events : (f == 2 ? &h->n_upd : &h->n_succ));
        const uint32_t cap = f == 0 ? G.xcap_q : (f == 1 ? G.xcap_ev : (f == 2 ? G.xcap_upd : G.xcap_succ));
        const uint32_t v = pull(p);
        raw[k][f] = v; cnt[k][f] = min(v, cap);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        uint32_t run = 0;
        for (int k = 0; k < nrk; k++) { off[k][threadIdx.x] = run; run += cnt[k][threadIdx.x]; }
        off[nrk][threadIdx.x] = run;
    }
    __syncthreads();
    const uint32_t qbase = c->n_q_base;
    const size_t qb = (size_t)(c->qsel ^ 1u) * G.cap_queue;
    {
        const int k = blockIdx.x % nrk;
        const uint32_t bt = (blockIdx.x / nrk) * blockDim.x + threadIdx.x, bsz = (gridDim.x / nrk) * blockDim.x;
        const XSlot x = xslot_of(G, k, day);
        const uint32_t n0 = cnt[k][0], n1 = n0 + cnt[k][1], n2 = n1 + (k != G.rank ? cnt[k][2] : 0u), n3 = n2 + cnt[k][3];
        for (uint32_t i = bt; i < n3; i += bsz) {
            if (i < n0) {                    // test-queue entries created by rank k's sweep
                const uint32_t d = qbase + off[k][0] + i;
                if (d < G.cap_queue) { G.q_key[qb + d] = pull(&x.q_key[i]); G.q_agent[qb + d] = pull(&x.q_agent[i]); }
            } else if (i < n1) {             // capacity events
                const uint32_t j = i - n0, d = off[k][1] + j;
                if (d < G.cap_events) { G.ev_key[d] = pull(&x.ev_key[j]); G.ev_agent[d] = pull(&x.ev_agent[j]); }
            } else if (i < n2) {             // state changes of rank k's agents (not the own ones: applied by the sweep itself)
                const uint2 u = pull(&x.upd[i - n1]); G.hot[u.x] = u.y;
            } else {                         // successful transmissions: first infector in sweep order wins
                const uint32_t j = i - n2, d = off[k][3] + j;
                if (d < G.cap_succ) { const Attempt at = pull(&x.succ[j]); G.succ[d] = at; atomicMin(&G.rec[at.cand].winner, at.key); }
            }
        }
    }
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (uint32_t i = gtid; i < RB_N_ATTRS * RB_MAX_AGES; i += gsz) {      // per-age counter deltas of the sweeps
        int d = 0;
        for (int k = 0; k < nrk; k++) d += pull(&xslot_of(G, k, day).hdr->counts[0][0] + i);
        if (d) (&c->counts[0][0])[i] += d;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x < RB_N_PLACES) {
        int d = 0;
        for (int k = 0; k < nrk; k++) d += pull(&xslot_of(G, k, day).hdr->daily_contacts[threadIdx.x]);
        c->daily_contacts[threadIdx.x] += d;
    }
    if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 64) {        // one warp: lane k reads rank k's scalars
        const int k = threadIdx.x - 32;
        int tor = 0, tio = 0, ex = 0, prob = 0;
        if (k < nrk) {
            const RepCtr *h = xslot_of(G, k, day).hdr;
            tor = pull(&h->total_infectors); tio = pull(&h->total_infections); ex = pull(&h->exposed_per_day); prob = pull(&h->problem);
            if (prob) set_problem(c, prob);
            if (raw[k][0] > G.xcap_q || raw[k][1] > G.xcap_ev || raw[k][2] > G.xcap_upd || raw[k][3] > G.xcap_succ) set_problem(c, RB_OTHER_FAILURE);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            tor += __shfl_xor_sync(0xffffffffu, tor, o); tio += __shfl_xor_sync(0xffffffffu, tio, o); ex += __shfl_xor_sync(0xffffffffu, ex, o);
        }
        if (k == 0) {
            c->total_infectors += tor; c->total_infections += tio; c->exposed_per_day += ex;
            if (qbase + off[nrk][0] > G.cap_queue || off[nrk][1] > G.cap_events || off[nrk][3] > G.cap_succ) set_problem(c, RB_OTHER_FAILURE);
            c->n_newq = min(qbase + off[nrk][0], G.cap_queue); c->n_events = min(off[nrk][1], G.cap_events); c->n_succ = min(off[nrk][3], G.cap_succ);
        }
    }
}

#endif
