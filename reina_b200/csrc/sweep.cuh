// reina_b200 / csrc / sweep.cuh
// k_sweep: Context._iterate_people / person_advance over the active agents (main.pyx:1968-1992, 395-438).
#ifndef REINA_B200_SWEEP_CUH
#define REINA_B200_SWEEP_CUH
#include "state.cuh"

// ---------------------------------------------------------------- k_sweep
// The daily sweep = Context._iterate_people / _process_person / person_advance (main.pyx:1968-1992, 395-438).
//
// Every warp streams its share of the packed words (coalesced 16-byte loads, 256 agents per step) and pushes the
// few agents that have anything to do today into a private shared-memory ring.  Work then flows through three
// warp-private rings, each drained only in full batches of 32 so that every stage executes on dense warps and no
// block-level barrier exists anywhere:
//   ring A (active agents)   -> stage 1: R bookkeeping, "infected today" flag, day counters, transition detection
//   ring E (infectious)      -> stage E: number of contacts (one Philox block + tabulated distribution), contact
//                               work items allocated with a warp prefix sum + one atomic and written coalesced
//   ring T (state changes)   -> stage T: symptom onset (gamma draw, durations, testing queue), end of illness,
//                               ward / ICU exits (capacity events tagged with the agent's sweep position)
#ifndef SW_THREADS
#define SW_THREADS 128
#endif
#define SW_WARPS (SW_THREADS / 32)
#define SW_CHUNK 256
#define SW_QCAP 256          // ring A takes at most 128 entries per step on top of < 32 left over
#define SW_RCAP 64
#ifndef SW_PFD
#define SW_PFD 3             // packed-word chunks in flight per warp (cp.async), 1 KB each; 0 = plain loads
#endif
#ifndef SW_CTAS_PER_SM
#define SW_CTAS_PER_SM 9     // 56 registers; measured 6 / 8 / 9 / 10 / 12 CTAs per SM: 170.0 / 162.7 / 159.0 / 161.0 / 165.7 ms per HUS step
#endif

struct WarpRings {
    uint32_t qi[SW_QCAP], qw[SW_QCAP];      // ring A: agent index, packed word as streamed (dense days)
    uint32_t ea[SW_RCAP], ed[SW_RCAP];      // ring E: agent index, contact descriptor
    uint32_t ta[SW_RCAP], tw[SW_RCAP];      // ring T: agent index, packed word (day counters already advanced)
};

// warp-aggregated push of (x, y) for the lanes with `want` into a ring of SW_RCAP entries; returns the new tail
__device__ __forceinline__ uint32_t ring_push(uint32_t *ra, uint32_t *rb, uint32_t tail, bool want, uint32_t x, uint32_t y, int lane) {
    const uint32_t m = __ballot_sync(0xffffffffu, want);
    if (want) { uint32_t p = (tail + __popc(m & ((1u << lane) - 1u))) & (SW_RCAP - 1); ra[p] = x; rb[p] = y; }
    return tail + __popc(m);
}

// Where the sweep puts what other kernels (and, in population-sharded mode, other ranks) consume.  Single GPU: the
// replica's own counters and lists.  Sharded: this rank's message slot, merged on every rank by k_merge.
struct SweepOut {
    RepCtr *cd;                                        // counters the sweep ADDS to
    unsigned long long *q_key; int32_t *q_agent; uint32_t cap_q;
    unsigned long long *ev_key; int32_t *ev_agent; uint32_t cap_ev;
    uint2 *upd; uint32_t cap_upd;                      // null on a single GPU
};
__device__ __forceinline__ SweepOut sweep_out(const Eng &G, int r, RepCtr *c) {
    SweepOut O;
    if (!G.xbuf) {
        const size_t qb = ((size_t)r * 2 + (c->qsel ^ 1u)) * G.cap_queue;
        O.cd = c; O.q_key = G.q_key + qb; O.q_agent = G.q_agent + qb; O.cap_q = G.cap_queue;
        O.ev_key = G.ev_key + (size_t)r * G.cap_events; O.ev_agent = G.ev_agent + (size_t)r * G.cap_events; O.cap_ev = G.cap_events;
        O.upd = nullptr; O.cap_upd = 0;
    } else {
        const XSlot x = xslot_of(G, G.rank, c->day);
        O.cd = x.hdr; O.q_key = x.q_key; O.q_agent = x.q_agent; O.cap_q = G.xcap_q;
        O.ev_key = x.ev_key; O.ev_agent = x.ev_agent; O.cap_ev = G.xcap_ev; O.upd = x.upd; O.cap_upd = G.xcap_upd;
    }
    return O;
}

// stage E: get_exposed_people / get_nr_contacts (main.pyx:936-955, 1308-1320) + work-item emission
__device__ __forceinline__ void stage_expose(const Eng &G, RepCtr *c, RepCtr *cd, const DevTable *tb, const WarpRings &W, uint32_t head, uint32_t m,
                                             uint2 *items, int lane) {
    uint32_t cnt = 0, ncont = 0, desc = 0, a = 0;
    if ((uint32_t)lane < m) {
        a = W.ea[(head + lane) & (SW_RCAP - 1)];
        desc = W.ed[(head + lane) & (SW_RCAP - 1)];
        const int age = age_of(G, (int32_t)a);
        const int cls = (desc >> 22) & 1u;
        u32x4 x = philox(c->seed, a, (uint32_t)c->day, PU_NCONTACT, 0);
        const double u = u01d(x.x, x.y);
        // n = first k with u < cdf[k] (k = limit if none); entries below nguide[u's top 8 bits] cannot match
        const double *cdf = tb->ncdf[age][cls];
        const int limit = cls ? 5 : 100;
        int k = tb->nguide[age][cls][x.x >> 24];
        while (k < limit && !(u < __ldg(&cdf[k]))) k++;
        ncont = (uint32_t)k;
        cnt = (ncont + 3u) >> 2;          // work items are groups of four contact slots (they share one Philox block)
        desc = (desc & ~(1u << 22)) | ((uint32_t)age << 7);
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const uint32_t wtot = __shfl_sync(0xffffffffu, incl, 31);
    if (wtot == 0) return;
    const uint32_t excl = incl - cnt;
    uint32_t gbase = 0;
    uint32_t ctot = ncont;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ctot += __shfl_xor_sync(0xffffffffu, ctot, o);
    if (lane == 31) { gbase = atomicAdd(&c->n_items, wtot); atomicAdd(&cd->exposed_per_day, (int)ctot); }
    gbase = __shfl_sync(0xffffffffu, gbase, 31);
    if (gbase + wtot > G.cap_items) { if (lane == 0) set_problem(cd, RB_OTHER_FAILURE); return; }
    for (uint32_t t0 = 0; t0 < wtot; t0 += 32) {
        const uint32_t t = t0 + lane;
        int lo = 0;     // owner = largest lane whose exclusive prefix is <= t
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            uint32_t e = __shfl_sync(0xffffffffu, excl, (lo + step) & 31);
            if (lo + step < 32 && e <= t) lo += step;
        }
        const uint32_t oa = __shfl_sync(0xffffffffu, a, lo), od = __shfl_sync(0xffffffffu, desc, lo), oe = __shfl_sync(0xffffffffu, excl, lo);
        const uint32_t on = __shfl_sync(0xffffffffu, ncont, lo);
        if (t < wtot) {
            const uint32_t g = t - oe, left = on - 4u * g;            // group index, contacts from this group on
            items[gbase + t] = make_uint2(oa, od | g | (((left < 4u ? left : 4u) - 1u) << 5));
        }
    }
}

__device__ __forceinline__ void emit_event(const Eng &G, const SweepOut &O, RepCtr *c, int32_t a, int type) {
    uint32_t idx = atomicAdd(&O.cd->n_events, 1u);
    if (idx < O.cap_ev) {
        O.ev_key[idx] = ((unsigned long long)sweep_pos(G, c, (uint32_t)a) << 2) | (unsigned)type;
        O.ev_agent[idx] = a;
    } else set_problem(O.cd, RB_OTHER_FAILURE);
}

// stage T: the state changes of person_advance (main.pyx:405-438) for agents whose day counter reached zero
__device__ __forceinline__ void stage_transition_lane(const Eng &G, int r, RepCtr *c, const SweepOut &O, const WarpRings &W, uint32_t head, int lane,
                                                      int32_t &a_out, uint32_t &h_out) {
    RepCtr *cd = O.cd;
    const size_t base = (size_t)r * G.Npad;
    const int32_t a = (int32_t)W.ta[(head + lane) & (SW_RCAP - 1)];
    uint32_t h = W.tw[(head + lane) & (SW_RCAP - 1)];
    const int day = c->day;
    const int age = age_of(G, a);
    const uint32_t st = H_STATE(h), sev = H_SEV(h);
    const rb_variant *v = &G.variants[H_VAR(h)];
    if (st == RB_INCUBATION) {
        // person_become_ill, main.pyx:284-291; durations :989-1039 fixed from the one onset-to-removed draw
        float T = (sev == RB_FATAL)
            ? gamma_f(c->seed, (uint32_t)a, (uint32_t)day, PU_ONSET, v->onset_death_kappa, v->onset_death_theta)
            : gamma_f(c->seed, (uint32_t)a, (uint32_t)day, PU_ONSET, v->onset_recovery_kappa, v->onset_recovery_theta);
        float f = T;
        if (sev != RB_ASYMPTOMATIC && sev != RB_MILD) f = f * v->ratio_before_hospitalisation;
        const uint32_t dl = (uint32_t)clamp255(round_to_int(f));
        float w = 0.0f, u = 0.0f;
        if (sev == RB_SEVERE) w = T * (1.0f - v->ratio_before_hospitalisation);
        else if (sev == RB_CRITICAL || sev == RB_FATAL) {
            w = T * v->ratio_in_ward;
            u = ((1.0f - v->ratio_in_ward) - v->ratio_before_hospitalisation) * T;
        }
        const uint32_t wd = (uint32_t)clamp255(round_to_int(w)), ud = (uint32_t)clamp255(round_to_int(u));
        if (wd | ud) atomicOr(&G.rec[base + a].cold, (wd << 16) | (ud << 24));
        h = H_SET_DL(H_SET_STATE(h, RB_ILLNESS), dl);
        if (sev != RB_ASYMPTOMATIC && !(h & H_DET)) {
            // seek_testing, main.pyx:595-615
            bool q = false;
            const int mode = c->testing_mode;
            if (mode == RB_ALL_WITH_SYMPTOMS || mode == RB_ALL_WITH_SYMPTOMS_CT) q = true;
            else if (mode == RB_ONLY_SEVERE_SYMPTOMS) {
                if (sev >= RB_SEVERE) q = true;
                else {
                    u32x4 x = philox(c->seed, (uint32_t)a, (uint32_t)day, PU_SEEK, 0);
                    q = chance(u01d(x.x, x.y), c->p_detected_anyway);
                }
            }
            if (q && !(h & H_QUEUED)) {     // queue_for_testing guards (not DEAD / detected / queued), main.pyx:476
                h |= H_QUEUED;
                uint32_t idx = atomicAdd(&cd->n_newq, 1u);
                if (idx < O.cap_q) {
                    O.q_key[idx] = QKEY_SWEEP | sweep_pos(G, c, (uint32_t)a);
                    O.q_agent[idx] = a;
                } else set_problem(cd, RB_OTHER_FAILURE);
            }
        }
    } else if (st == RB_ILLNESS) {
        if (sev == RB_FATAL) {                       // person_die, main.pyx:370-374, 1618-1623
            h = H_SET_STATE(h, RB_DEAD) & ~H_LIST;
            count_add(cd, RB_A_INFECTED, age, -1); count_add(cd, RB_A_DEAD, age, 1); count_add(cd, RB_A_NON_HOSPITAL_DEATHS, age, 1);
        } else if (sev >= RB_SEVERE) {               // person_hospitalize, main.pyx:321-338: the bed claim is an event
            if (!(h & H_DET)) { h |= H_DET; count_add(cd, RB_A_DETECTED, age, 1); count_add(cd, RB_A_ALL_DETECTED, age, 1); }
            emit_event(G, O, c, a, EV_HOSP_CLAIM);
        } else {                                     // person_recover, main.pyx:315-318
            h = H_SET_STATE(h, RB_RECOVERED) & ~H_LIST;
            count_add(cd, RB_A_INFECTED, age, -1); count_add(cd, RB_A_RECOVERED, age, 1);
        }
    } else {   // HOSPITALIZED / IN_ICU
        int type;
        if (st == RB_HOSPITALIZED && (sev == RB_CRITICAL || sev == RB_FATAL)) type = EV_TO_ICU;   // main.pyx:430-431
        else {
            // person_release_from_hospital, main.pyx:354-367: the outcome does not depend on capacity
            type = st == RB_IN_ICU ? EV_ICU_RELEASE : EV_WARD_RELEASE;
            count_add(cd, st == RB_IN_ICU ? RB_A_IN_ICU : RB_A_IN_WARD, age, -1);
            count_add(cd, RB_A_INFECTED, age, -1);
            if (sev == RB_FATAL) { h = H_SET_STATE(h, RB_DEAD) & ~H_LIST; count_add(cd, RB_A_DEAD, age, 1); count_add(cd, RB_A_NON_HOSPITAL_DEATHS, age, 1); }
            else { h = H_SET_STATE(h, RB_RECOVERED) & ~H_LIST; count_add(cd, RB_A_RECOVERED, age, 1); }
        }
        emit_event(G, O, c, a, type);
    }
    G.hot[base + a] = h;
    a_out = a; h_out = h;
}
__device__ __forceinline__ void stage_transition(const Eng &G, int r, RepCtr *c, const WarpRings &W, uint32_t head, uint32_t m, int lane) {
    const SweepOut O = sweep_out(G, r, c);     // resolved here, not in the caller: the streaming loop stays light on registers
    int32_t a = 0; uint32_t h = 0;
    const bool on = (uint32_t)lane < m;
    if (on) stage_transition_lane(G, r, c, O, W, head, lane, a, h);
    if (O.upd) {      // sharded mode: the other ranks' copies of this agent learn the new state and flags from the log
        const uint32_t mk = __ballot_sync(0xffffffffu, on);
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(&O.cd->n_upd, (uint32_t)__popc(mk));
        b = __shfl_sync(0xffffffffu, b, 0);
        if (on) {
            const uint32_t idx = b + __popc(mk & ((1u << lane) - 1u));
            if (idx < O.cap_upd) O.upd[idx] = make_uint2((uint32_t)a, h); else set_problem(O.cd, RB_OTHER_FAILURE);
        }
    }
}

// stage 1 for one active agent: R bookkeeping, "infected today" flag, day counters, what happens next
__device__ __forceinline__ void stage_active_lane(const Eng &G, int r, RepCtr *c, size_t base, uint32_t a, uint32_t &h, bool &want_e, bool &want_t,
                                                  bool &removed, int &infected_others, uint32_t &desc) {
    const uint32_t st = H_STATE(h);
    if (st >= RB_RECOVERED) {          // R bookkeeping, main.pyx:1969-1972 (only agents not yet included reach here)
        removed = true;
        infected_others = (int)(G.rec[base + a].cold & 0xffffu);
        G.hot[base + a] = h | H_INCL;
        atomicAnd(&G.act[(size_t)r * G.sus_words + (a >> 5)], ~(1u << (a & 31)));   // nothing left to do for this agent
    } else if (h & H_FRESH) {          // infected today before the sweep: wait until tomorrow, main.pyx:402-403
        G.hot[base + a] = h & ~H_FRESH;
    } else {
        const uint32_t sev = H_SEV(h), var = H_VAR(h);
        uint32_t dl = H_DL(h);
        if (st == RB_INCUBATION || st == RB_ILLNESS) {
            const int dayidx = st == RB_INCUBATION ? -(int)dl : (int)H_DOI(h);
            if (!(h & H_DET) && dayidx >= -10 && dayidx <= 10 && G.variants[var].iot[dayidx + 10] != 0.0f) {
                want_e = true;
                const uint32_t cls = (st == RB_ILLNESS && sev != RB_ASYMPTOMATIC) ? 1u : 0u;   // factor 0.5, limit 5
                desc = ((uint32_t)(dayidx + 10) << 14) | ((sev == RB_ASYMPTOMATIC ? 1u : 0u) << 19) | (var << 20) | (cls << 22);
            }
            if (st == RB_ILLNESS) { uint32_t doi = H_DOI(h); if (doi < 31) doi++; h = H_SET_DOI(h, doi); }
        }
        if (dl > 0) dl--;
        h = H_SET_DL(h, dl);
        if (dl == 0) want_t = true; else G.hot[base + a] = h;
    }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int nbytes = valid ? 16 : 0;      // src-size 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// the four packed words of one lane -> ring A (index and word); returns the new tail
__device__ __forceinline__ uint32_t sweep_push4(WarpRings &W, uint32_t tail, const uint4 w, uint32_t act, uint32_t a_first, int lane) {
    const uint32_t hw[4] = {w.x, w.y, w.z, w.w};
    if (!__any_sync(0xffffffffu, act != 0)) return tail;
    const uint32_t mine = __popc(act);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t p = tail + incl - mine;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (act & (1u << j)) { W.qi[p & (SW_QCAP - 1)] = a_first + j; W.qw[p & (SW_QCAP - 1)] = hw[j]; p++; }
    return tail + tot;
}

// The kernel is one producer / consumer loop per warp.  The producer fills ring A from today's source -- the packed
// words themselves on dense days, the activity bitmap on sparse days -- until a full batch of 32 is queued; the
// consumer runs each stage on one batch.  Every stage is instantiated exactly ONCE: the stages are thousands of
// instructions each, and a second inlined copy in the hot loop pushes it out of the instruction cache.
__global__ void __launch_bounds__(SW_THREADS, SW_CTAS_PER_SM) k_sweep(Eng G) {
    __shared__ WarpRings s_rings[SW_WARPS];
#if SW_PFD > 0
    __shared__ uint4 s_pf[SW_WARPS][SW_PFD][2][32];
#endif
    const int r = blockIdx.y + G.r0;
    RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const DevTable *tb = G.tables[c->epoch];
    uint2 *items = G.items + (size_t)r * G.cap_items;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpRings &W = s_rings[warp];
    RepCtr *cd = !G.xbuf ? c : xslot_of(G, G.rank, c->day).hdr;      // counters the sweep adds to
    const int nrk = G.nranks, rk = G.rank;
    const int stride = gridDim.x * SW_WARPS;
    const bool stream = c->stream_mode != 0;
    uint32_t head = 0, tail = 0, e_head = 0, e_tail = 0, t_head = 0, t_tail = 0;
    uint32_t ready = 0;                     // bitmap walk: ring A entries below `ready` were committed before the latest gather group

    // ---- producer state.  Dense day: the packed words are streamed, 256 agents (1 KB) per warp step, in two halves of
    // 128 so that ring A never takes more than 128 entries at once.  Sparse day: one bit per agent says whether the
    // sweep has anything to do for it (infected, or removed and not yet counted in R), so the pass over all N agents
    // reads 1/32 of the packed state -- an L2-resident bitmap -- and only the active agents' words are gathered; a
    // warp step covers 32 lanes x 128 agents = one ownership stripe.
    const uint4 *hot4 = reinterpret_cast<const uint4 *>(G.hot + base);
    const uint4 *act4 = reinterpret_cast<const uint4 *>(G.act + (size_t)r * G.sus_words);
    const int n_chunks = (G.Npad + SW_CHUNK - 1) / SW_CHUNK, n4 = G.Npad >> 2, n_vec = G.sus_words >> 2;
    const int n_mine = stream ? (((n_chunks + 15) >> 4) + nrk - 1) / nrk * 16        // this rank's chunks: 16 per stripe
                              : (((n_vec + 31) >> 5) + nrk - 1) / nrk;              // this rank's bitmap steps
    int j = blockIdx.x * SW_WARPS + warp;
    bool more = j < n_mine;
    uint4 wb = make_uint4(0, 0, 0, 0);      // dense: second half of the current chunk; sparse: this lane's 128 activity bits
    uint32_t cw = 0, a0 = 0, abits = 0;
    int part = 0;                           // dense: 0 = load a chunk, 1 = second half pending; sparse: word of `wb` in `cw` (4 = none)
    if (!stream) part = 4;
    // bitmap walk: the activity vector of the NEXT warp step is loaded one step ahead
    auto load_act = [&](int jj) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (jj < n_mine) { const int vi = (nrk == 1 ? jj : jj * nrk + rk) * 32 + lane; if (vi < n_vec) v = __ldg(&act4[vi]); }
        return v;
    };
    uint4 nb = make_uint4(0, 0, 0, 0);
    if (!stream) nb = load_act(j);
#if SW_PFD > 0
    uint4 (*pf)[2][32] = s_pf[warp];
    int slot = 0;
    auto chunk_of = [&](int jj) { return nrk == 1 ? jj : ((((jj >> 4) * nrk + rk) << 4) | (jj & 15)); };
    auto fetch = [&](int jj, int sl) {      // each lane copies its two 16-byte pieces of chunk jj into its own slots
        const int chunk = chunk_of(jj);
        const int i0 = chunk * (SW_CHUNK / 4) + lane, i1 = i0 + 32;
        const bool in = jj < n_mine && chunk < n_chunks;
        const bool v0 = in && i0 < n4, v1 = in && i1 < n4;
        cp_async16(&pf[sl][0][lane], hot4 + (v0 ? i0 : 0), v0);
        cp_async16(&pf[sl][1][lane], hot4 + (v1 ? i1 : 0), v1);
        cp_async_commit();
    };
    if (stream) {
#pragma unroll
        for (int d = 0; d < SW_PFD; d++) fetch(j + d * stride, d);
    }
#else
    auto chunk_of = [&](int jj) { return nrk == 1 ? jj : ((((jj >> 4) * nrk + rk) << 4) | (jj & 15)); };
#endif

    for (;;) {
        // ---------------- produce
        while (more && (stream ? tail : ready) - head < 32) {
            if (stream) {
                if (part == 0) {
                    const int chunk = chunk_of(j);
                    uint4 w0 = make_uint4(0, 0, 0, 0);
#if SW_PFD > 0
                    cp_async_wait<SW_PFD - 1>();
                    w0 = pf[slot][0][lane]; wb = pf[slot][1][lane];
                    fetch(j + SW_PFD * stride, slot);          // refill the slot just read
                    slot = slot + 1 == SW_PFD ? 0 : slot + 1;
#else
                    const int i0 = chunk * (SW_CHUNK / 4) + lane, i1 = i0 + 32;
                    wb = w0;
                    if (chunk < n_chunks && i0 < n4) w0 = hot4[i0];
                    if (chunk < n_chunks && i1 < n4) wb = hot4[i1];
#endif
                    a0 = (uint32_t)chunk * SW_CHUNK;
                    // who is active comes from the bitmap (2 words per lane out of the chunk's 8, one 32-byte sector per
                    // warp) rather than from decoding all eight packed words: this lane's agents are two nibbles
                    const uint32_t *aw = G.act + (size_t)r * G.sus_words + (a0 >> 5) + (lane >> 3);
                    uint32_t b0 = 0, b1 = 0;
                    if (chunk < n_chunks) { b0 = __ldg(aw); b1 = __ldg(aw + 4); }
                    b0 = (b0 >> ((lane & 7) * 4)) & 15u; abits = (b1 >> ((lane & 7) * 4)) & 15u;
                    tail = sweep_push4(W, tail, w0, b0, a0 + lane * 4, lane);
                    part = 1;
                } else {
                    tail = sweep_push4(W, tail, wb, abits, a0 + 128 + lane * 4, lane);
                    part = 0;
                    j += stride; more = j < n_mine;
                }
            } else {
                if (part == 4) {                               // next 32 x 128 activity bits
                    const int v0 = (nrk == 1 ? j : j * nrk + rk) * 32;
                    j += stride;
                    const int vi = v0 + lane;
                    wb = nb;
                    nb = load_act(j);
                    a0 = (uint32_t)vi * 128u;
                    if (__any_sync(0xffffffffu, (wb.x | wb.y | wb.z | wb.w) != 0u)) { part = 0; cw = wb.x; }
                    else more = j < n_mine;
                } else if (!__any_sync(0xffffffffu, cw != 0u)) {
                    part++;
                    cw = part == 1 ? wb.y : (part == 2 ? wb.z : wb.w);
                    if (part == 4) more = j < n_mine;
                } else if (tail - head > SW_QCAP - 128) {
                    // no room for another round of up to 128 entries: everything queued becomes consumable
                    cp_async_wait<0>();
                    ready = tail;
                } else {
                    // every lane queues up to 4 of its set bits per round: at most 128 pushes, the ring holds 256
                    const uint32_t mine = min(__popc(cw), 4);
                    uint32_t incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                    uint32_t p = tail + incl - mine;
                    // the active agents' packed words are gathered HERE (the only per-agent gather of the sweep), as
                    // asynchronous 4-byte copies straight into the ring: the producer never waits for them, the consumer
                    // waits for all but the latest round's group, so the gathers of several rounds are in flight while
                    // earlier batches run through the stages
#pragma unroll
                    for (uint32_t k = 0; k < 4; k++)
                        if (k < mine) {
                            const uint32_t ia = a0 + (uint32_t)part * 32u + (uint32_t)(__ffs(cw) - 1); cw &= cw - 1u;
                            W.qi[(p + k) & (SW_QCAP - 1)] = ia;
                            cp_async4(&W.qw[(p + k) & (SW_QCAP - 1)], G.hot + base + ia);
                        }
                    cp_async_commit();
                    ready = tail;
                    tail += tot;
                }
            }
            __syncwarp();
        }
        // ---------------- consume: full batches while the source lasts, whatever is left afterwards
        if (!stream) {
            if (more) cp_async_wait<1>(); else { cp_async_wait<0>(); ready = tail; }
            __syncwarp();
        }
        const uint32_t av = (stream ? tail : ready) - head;
        if (av) {
            const uint32_t m = min(32u, av);
            const size_t gb = base;
            bool want_e = false, want_t = false, removed = false;
            int infected_others = 0;
            uint32_t a = 0, h = 0, desc = 0;
            if ((uint32_t)lane < m) {
                a = W.qi[(head + lane) & (SW_QCAP - 1)];
                h = W.qw[(head + lane) & (SW_QCAP - 1)];      // gathered (bitmap walk) or streamed by the producer
                stage_active_lane(G, r, c, gb, a, h, want_e, want_t, removed, infected_others, desc);
            }
            head += m;
            const uint32_t rm = __ballot_sync(0xffffffffu, removed);
            if (rm) {                              // one pair of atomics per warp batch instead of one per removed agent
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) infected_others += __shfl_xor_sync(0xffffffffu, infected_others, o);
                if (lane == 0) { atomicAdd(&cd->total_infectors, __popc(rm)); if (infected_others) atomicAdd(&cd->total_infections, infected_others); }
            }
            e_tail = ring_push(W.ea, W.ed, e_tail, want_e, a, desc, lane);
            t_tail = ring_push(W.ta, W.tw, t_tail, want_t, a, h, lane);
            __syncwarp();
        }
        const bool flush = !more && tail == head;
        const uint32_t e_av = e_tail - e_head, t_av = t_tail - t_head;
        if (e_av >= 32 || (flush && e_av)) { const uint32_t m = min(32u, e_av); stage_expose(G, c, cd, tb, W, e_head, m, items, lane); e_head += m; }
        if (t_av >= 32 || (flush && t_av)) { const uint32_t m = min(32u, t_av); stage_transition(G, r, c, W, t_head, m, lane); t_head += m; }
        __syncwarp();
        if (flush && e_tail == e_head && t_tail == t_head) break;
    }
#if SW_PFD > 0
    cp_async_wait<0>();
#endif
}

#endif
