// reina_b200 / csrc / shard.cuh
// k_merge: population-sharded mode, applies every rank's message of the day (see include/reina_b200.h, rb_shard_init).
#ifndef REINA_B200_SHARD_CUH
#define REINA_B200_SHARD_CUH
#include "state.cuh"

// ---------------------------------------------------------------- k_merge (population-sharded mode only)
// After the all-gather every rank holds every rank's message.  All ranks apply all of them in rank order, so the
// replicated state (counters, test queue, capacity events, packed words, conflict slots) stays identical everywhere:
// count deltas are added, queue entries / events / successful transmissions are concatenated into the single-GPU
// lists, the other ranks' state changes overwrite the local copies of their agents, and every successful
// transmission does its atomicMin on the target's conflict slot (first infector in sweep order wins, main.pyx:238-244).
#define MAX_RANKS 16
__global__ void __launch_bounds__(256) k_merge(Eng G) {
    __shared__ uint32_t nq[MAX_RANKS + 1], ne[MAX_RANKS + 1], nu[MAX_RANKS + 1], ns[MAX_RANKS + 1];
    RepCtr *c = &G.ctr[0];
    const int nrk = G.nranks;
    if (threadIdx.x == 0) {
        uint32_t q = 0, e = 0, u = 0, sx = 0;
        for (int k = 0; k < nrk; k++) {
            const RepCtr *h = xslot_of(G, k).hdr;
            nq[k] = q; ne[k] = e; nu[k] = u; ns[k] = sx;
            q += min(h->n_newq, G.xcap_q); e += min(h->n_events, G.xcap_ev); u += min(h->n_upd, G.xcap_upd); sx += min(h->n_succ, G.xcap_succ);
        }
        nq[nrk] = q; ne[nrk] = e; nu[nrk] = u; ns[nrk] = sx;
    }
    __syncthreads();
    const uint32_t qbase = c->n_q_base;
    const size_t qb = (size_t)(c->qsel ^ 1u) * G.cap_queue;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int k = 0; k < nrk; k++) {
        const XSlot x = xslot_of(G, k);
        for (uint32_t i = gtid; i < nq[k + 1] - nq[k]; i += gsz) {
            const uint32_t d = qbase + nq[k] + i;
            if (d < G.cap_queue) { G.q_key[qb + d] = x.q_key[i]; G.q_agent[qb + d] = x.q_agent[i]; }
        }
        for (uint32_t i = gtid; i < ne[k + 1] - ne[k]; i += gsz) {
            const uint32_t d = ne[k] + i;
            if (d < G.cap_events) { G.ev_key[d] = x.ev_key[i]; G.ev_agent[d] = x.ev_agent[i]; }
        }
        if (k != G.rank)
            for (uint32_t i = gtid; i < nu[k + 1] - nu[k]; i += gsz) { const uint2 u = x.upd[i]; G.hot[u.x] = u.y; }
        for (uint32_t i = gtid; i < ns[k + 1] - ns[k]; i += gsz) {
            const uint32_t d = ns[k] + i;
            if (d < G.cap_succ) { const Attempt at = x.succ[i]; G.succ[d] = at; atomicMin(&G.rec[at.cand].winner, at.key); }
        }
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < RB_N_ATTRS * RB_MAX_AGES; i += blockDim.x) {
            int d = 0;
            for (int k = 0; k < nrk; k++) d += (&xslot_of(G, k).hdr->counts[0][0])[i];
            if (d) (&c->counts[0][0])[i] += d;
        }
        if (threadIdx.x < RB_N_PLACES) { int d = 0; for (int k = 0; k < nrk; k++) d += xslot_of(G, k).hdr->daily_contacts[threadIdx.x]; c->daily_contacts[threadIdx.x] += d; }
        if (threadIdx.x == 32) {
            for (int k = 0; k < nrk; k++) {
                const RepCtr *h = xslot_of(G, k).hdr;
                c->total_infectors += h->total_infectors; c->total_infections += h->total_infections; c->exposed_per_day += h->exposed_per_day;
                if (h->problem) set_problem(c, h->problem);
                if (h->n_newq > G.xcap_q || h->n_events > G.xcap_ev || h->n_upd > G.xcap_upd || h->n_succ > G.xcap_succ) set_problem(c, RB_OTHER_FAILURE);
            }
            if (qbase + nq[nrk] > G.cap_queue || ne[nrk] > G.cap_events || ns[nrk] > G.cap_succ) set_problem(c, RB_OTHER_FAILURE);
            c->n_newq = min(qbase + nq[nrk], G.cap_queue); c->n_events = min(ne[nrk], G.cap_events); c->n_succ = min(ns[nrk], G.cap_succ);
        }
    }
}

#endif
