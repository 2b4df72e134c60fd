// reina_b200 / csrc / shard.cuh
// k_merge: population-sharded mode, applies every rank's message of the day (see include/reina_b200.h, rb_shard_init).
#ifndef REINA_B200_SHARD_CUH
#define REINA_B200_SHARD_CUH
#include "state.cuh"

// ---------------------------------------------------------------- peer exchange: publish / wait
// Exchange through peer memory (Eng::xp2p): once its sweep and contact kernels are done, a rank raises the flag at the
// head of its own buffer to xflag_value(day); k_merge on every rank waits for every owner's flag and then reads the
// message straight out of the owner's memory over NVLink -- header first, then exactly as many list entries as the
// header counts -- instead of receiving fixed-size slots from an all-gather.  The slots alternate with the day's
// parity: a rank overwrites the slot of day d on day d+2, after its day-d+1 merge, which waited for every peer's
// day-d+1 flag, which every peer raised after finishing its day-d merge.
__global__ void k_publish(Eng G) {
    if (threadIdx.x != 0) return;
    __threadfence_system();                                   // the kernels before this one wrote the message
    *(volatile uint32_t *)G.xpeer[G.rank] = xflag_value(G, G.ctr[0].day);
}
__device__ __forceinline__ void wait_for_peers(const Eng &G, RepCtr *c) {
    if ((int)threadIdx.x < G.nranks) {
        const volatile uint32_t *flag = (const volatile uint32_t *)G.xpeer[threadIdx.x];
        const uint32_t want = xflag_value(G, c->day);
        const long long t0 = clock64();
        while ((int32_t)(*flag - want) < 0)
            if (clock64() - t0 > 20000000000ll) { set_problem(c, RB_OTHER_FAILURE); break; }     // ~10 s: a peer died; fail loudly, do not hang
        __threadfence_system();
    }
    __syncthreads();
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
__global__ void __launch_bounds__(256) k_merge(Eng G) {
    __shared__ uint32_t nq[MAX_RANKS + 1], ne[MAX_RANKS + 1], nu[MAX_RANKS + 1], ns[MAX_RANKS + 1];
    RepCtr *c = &G.ctr[0];
    const int nrk = G.nranks;
    const int day = c->day;
    if (G.xp2p) wait_for_peers(G, c);
    if (threadIdx.x == 0) {
        uint32_t q = 0, e = 0, u = 0, sx = 0;
        for (int k = 0; k < nrk; k++) {
            const RepCtr *h = xslot_of(G, k, day).hdr;
            nq[k] = q; ne[k] = e; nu[k] = u; ns[k] = sx;
            q += min(pull(&h->n_newq), G.xcap_q); e += min(pull(&h->n_events), G.xcap_ev); u += min(pull(&h->n_upd), G.xcap_upd); sx += min(pull(&h->n_succ), G.xcap_succ);
        }
        nq[nrk] = q; ne[nrk] = e; nu[nrk] = u; ns[nrk] = sx;
    }
    __syncthreads();
    const uint32_t qbase = c->n_q_base;
    const size_t qb = (size_t)(c->qsel ^ 1u) * G.cap_queue;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int k = 0; k < nrk; k++) {
        const XSlot x = xslot_of(G, k, day);
        for (uint32_t i = gtid; i < nq[k + 1] - nq[k]; i += gsz) {
            const uint32_t d = qbase + nq[k] + i;
            if (d < G.cap_queue) { G.q_key[qb + d] = pull(&x.q_key[i]); G.q_agent[qb + d] = pull(&x.q_agent[i]); }
        }
        for (uint32_t i = gtid; i < ne[k + 1] - ne[k]; i += gsz) {
            const uint32_t d = ne[k] + i;
            if (d < G.cap_events) { G.ev_key[d] = pull(&x.ev_key[i]); G.ev_agent[d] = pull(&x.ev_agent[i]); }
        }
        if (k != G.rank)
            for (uint32_t i = gtid; i < nu[k + 1] - nu[k]; i += gsz) { const uint2 u = pull(&x.upd[i]); G.hot[u.x] = u.y; }
        for (uint32_t i = gtid; i < ns[k + 1] - ns[k]; i += gsz) {
            const uint32_t d = ns[k] + i;
            if (d < G.cap_succ) { const Attempt at = pull(&x.succ[i]); G.succ[d] = at; atomicMin(&G.rec[at.cand].winner, at.key); }
        }
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < RB_N_ATTRS * RB_MAX_AGES; i += blockDim.x) {
            int d = 0;
            for (int k = 0; k < nrk; k++) d += pull(&xslot_of(G, k, day).hdr->counts[0][0] + i);
            if (d) (&c->counts[0][0])[i] += d;
        }
        if (threadIdx.x < RB_N_PLACES) { int d = 0; for (int k = 0; k < nrk; k++) d += pull(&xslot_of(G, k, day).hdr->daily_contacts[threadIdx.x]); c->daily_contacts[threadIdx.x] += d; }
        if (threadIdx.x == 32) {
            for (int k = 0; k < nrk; k++) {
                const RepCtr *h = xslot_of(G, k, day).hdr;
                c->total_infectors += pull(&h->total_infectors); c->total_infections += pull(&h->total_infections); c->exposed_per_day += pull(&h->exposed_per_day);
                const int hp = pull(&h->problem);
                if (hp) set_problem(c, hp);
                if (pull(&h->n_newq) > G.xcap_q || pull(&h->n_events) > G.xcap_ev || pull(&h->n_upd) > G.xcap_upd || pull(&h->n_succ) > G.xcap_succ) set_problem(c, RB_OTHER_FAILURE);
            }
            if (qbase + nq[nrk] > G.cap_queue || ne[nrk] > G.cap_events || ns[nrk] > G.cap_succ) set_problem(c, RB_OTHER_FAILURE);
            c->n_newq = min(qbase + nq[nrk], G.cap_queue); c->n_events = min(ne[nrk], G.cap_events); c->n_succ = min(ns[nrk], G.cap_succ);
        }
    }
}

#endif
