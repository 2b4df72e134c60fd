// reina_b200 / csrc / sweep.cuh
// k_sweep: Context._iterate_people / person_advance over the active agents (main.pyx:1968-1992, 395-438).
#ifndef REINA_B200_SWEEP_CUH
#define REINA_B200_SWEEP_CUH
#include "state.cuh"

// ---------------------------------------------------------------- k_sweep
// The daily sweep = Context._iterate_people / _process_person / person_advance (main.pyx:1968-1992, 395-438).
//
// The reference walks all N agents every day to find the few per cent that are infected.  Here every replica keeps
// dense ACTIVE LISTS: one 8-byte entry (agent, copy of its packed word with today's day counters) per agent that is
// infected, or removed and not yet counted in R, in n_seg segments (agent a lives in segment a % n_seg for good).  A warp
// of the sweep owns whole segments: it streams segment s of today's buffer coalesced, advances the day counters in the
// copies and writes the survivors, compacted, to segment s of tomorrow's buffer behind a counter it keeps in a register
// -- no atomics, no other warp ever touches that segment during the sweep.  `hot`, the authoritative packed word that
// every other kernel gathers, is touched only when an agent changes state.  New infections are appended by k_resolve /
// the imports (device_infect -> list_add).  The order inside a segment is arbitrary and differs from run to run;
// nothing observable depends on it, because every order-dependent step of the reference is keyed on the agent's sweep
// position and every draw on (seed, agent, day, purpose).
//
// Work flows through two warp-private shared-memory rings, each drained only in full batches of 32, so the two heavy
// stages execute on dense warps and no block-level barrier exists anywhere:
//   stage 1 (every entry)     day counters, is the agent infectious today, is anything else due; survivors -> front of
//                             tomorrow's segment
//   ring E (infectious)    -> stage E: number of contacts (one Philox block + tabulated distribution), contact work
//                             items reserved with one scan + one atomic per batch
//   ring T (something due) -> stage T: everything that needs the authoritative word or the agent record: state changes
//                             (symptom onset with its gamma draw, end of illness, ward / ICU exits -> capacity events
//                             tagged with the sweep position), pending bed / ICU claims decided by the day boundary,
//                             R bookkeeping of removed agents; what stays listed -> back of tomorrow's segment
#ifndef SW_THREADS
#define SW_THREADS 128
#endif
#define SW_WARPS (SW_THREADS / 32)
#define SW_RCAP 64
#ifndef SW_CTAS_PER_SM
#define SW_CTAS_PER_SM 9      // 56 registers, no spills (10 CTAs = 48 registers spills the stage state)
#endif

struct WarpRings {
    uint32_t ea[SW_RCAP], ed[SW_RCAP];      // ring E: agent index, contact descriptor
    uint32_t ta[SW_RCAP], tw[SW_RCAP];      // ring T: agent index, list copy of the packed word (day counters already advanced)
};

// warp-aggregated push of (x, y) for the lanes with `want` into a ring of SW_RCAP entries; returns the new tail
__device__ __forceinline__ uint32_t ring_push(uint32_t *ra, uint32_t *rb, uint32_t tail, bool want, uint32_t x, uint32_t y, int lane) {
    const uint32_t m = __ballot_sync(0xffffffffu, want);
    if (want) { uint32_t p = (tail + __popc(m & ((1u << lane) - 1u))) & (SW_RCAP - 1); ra[p] = x; rb[p] = y; }
    return tail + __popc(m);
}

// The lanes with `want` append (a, w) to the segment this warp owns: `out_n` is the warp's private entry counter.
__device__ __forceinline__ void seg_append(const Eng &G, RepCtr *c, uint2 *out, uint32_t &out_n, bool want, uint32_t a, uint32_t w, int lane) {
    const uint32_t m = __ballot_sync(0xffffffffu, want);
    if (want) {
        const uint32_t p = out_n + __popc(m & ((1u << lane) - 1u));
        if (p < G.seg_cap) out[p] = make_uint2(a, w); else set_problem(c, RB_OTHER_FAILURE);
    }
    out_n += __popc(m);
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
__device__ __forceinline__ void stage_expose(const Eng &G, int r, RepCtr *c, RepCtr *cd, const DevTable *tb, const WarpRings &W, uint32_t head, uint32_t m,
                                             uint2 *items, int lane) {
    uint32_t ncont = 0, desc = 0, a = 0;
    if ((uint32_t)lane < m) {
        a = W.ea[(head + lane) & (SW_RCAP - 1)];
        desc = W.ed[(head + lane) & (SW_RCAP - 1)];
        // A detection the list copy could not know of (a traced contact, detected by this morning's queue drain): a
        // detected agent exposes nobody (person_expose_others, main.pyx:247-249).  Only consulted once tracing has been on.
        bool detected = false;
        if (c->ct_ever) detected = (__ldg(&G.det[(size_t)r * G.sus_words + (a >> 5)]) >> (a & 31)) & 1u;
        if (!detected) {
            const int age = age_of(G, (int32_t)a);
            const int cls = (desc >> 22) & 1u;
            u32x4 x = philox(c->seed, a, (uint32_t)c->day, PU_NCONTACT, 0);
            const double u = u01d(x.x, x.y);
            // n = first k with u < cdf[k] (k = limit if none); nguide[u's top 8 bits] = the answer for most cells, else where to start
            const double *cdf = tb->ncdf[age][cls];
            const int limit = cls ? 5 : 100;
            const uint32_t ng = tb->nguide[age][cls][x.x >> 24];
            int k = (int)(ng & 127u);
            if (ng & 128u) while (k < limit && !(u < __ldg(&cdf[k]))) k++;
            ncont = (uint32_t)k;
            desc = (desc & ~(1u << 22)) | ((uint32_t)age << 7);
        }
    }
    // work items are groups of four contact slots (they share one Philox block).  One scan for both totals: items in
    // the low half (<= 32 x 25), contacts in the high half (<= 32 x 100)
    const uint32_t cnt = (ncont + 3u) >> 2;
    uint32_t incl = cnt | (ncont << 16);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31), wtot = tot & 0xffffu;
    if (wtot == 0) return;
    uint32_t gbase = 0;
    if (lane == 31) { gbase = atomicAdd(&c->n_items, wtot); atomicAdd(&cd->exposed_per_day, (int)(tot >> 16)); }
    gbase = __shfl_sync(0xffffffffu, gbase, 31);
    if (gbase + wtot > G.cap_items) { if (lane == 0) set_problem(cd, RB_OTHER_FAILURE); return; }
    // every lane writes the items of its own agent (1-2 for most, 25 at the cap): the warp runs as long as its busiest lane
    uint2 *mine = items + gbase + ((incl & 0xffffu) - cnt);
    for (uint32_t g = 0; g < cnt; g++) {
        const uint32_t left = ncont - 4u * g;            // contacts from this group on
        mine[g] = make_uint2(a, desc | g | (((left < 4u ? left : 4u) - 1u) << 5));
    }
}

__device__ __forceinline__ void emit_event(const Eng &G, int r, const SweepOut &O, RepCtr *c, int32_t a, int type) {
    uint32_t idx = atomicAdd(&O.cd->n_events, 1u);
    if (idx < O.cap_ev) {
        O.ev_key[idx] = ((unsigned long long)sweep_pos(G, r, c, (uint32_t)a) << 2) | (unsigned)type;
        O.ev_agent[idx] = a;
    } else set_problem(O.cd, RB_OTHER_FAILURE);
}

// The state changes of person_advance (main.pyx:405-438) for an agent whose day counter reached zero.  `h` is the
// authoritative word with current day counters; returns the new word (written to `hot` by the caller) and, in `lw`, the
// copy tomorrow's sweep reads (H_PEND: a capacity claim is pending; H_DET preset: the agent joined the test queue and
// will be detected by tomorrow morning's drain before that sweep runs, main.pyx:514-545).
__device__ __forceinline__ uint32_t transition(const Eng &G, int r, RepCtr *c, const SweepOut &O, int32_t a, uint32_t h, uint32_t &lw) {
    RepCtr *cd = O.cd;
    const size_t base = (size_t)r * G.Npad;
    const int day = c->day;
    const int age = age_of(G, a);
    const uint32_t st = H_STATE(h), sev = H_SEV(h);
    const rb_variant *v = &G.variants[H_VAR(h)];
    uint32_t extra = 0;
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
                    O.q_key[idx] = QKEY_SWEEP | sweep_pos(G, r, c, (uint32_t)a);
                    O.q_agent[idx] = a;
                } else set_problem(cd, RB_OTHER_FAILURE);
            }
        }
        if (h & H_QUEUED) extra = H_DET;      // queued by this sweep or, earlier today, by contact tracing: detected tomorrow morning
    } else if (st == RB_ILLNESS) {
        if (sev == RB_FATAL) {                       // person_die, main.pyx:370-374, 1618-1623
            h = H_SET_STATE(h, RB_DEAD) & ~H_LIST;
            count_add(cd, RB_A_INFECTED, age, -1); count_add(cd, RB_A_DEAD, age, 1); count_add(cd, RB_A_NON_HOSPITAL_DEATHS, age, 1);
        } else if (sev >= RB_SEVERE) {               // person_hospitalize, main.pyx:321-338: the bed claim is an event
            if (!(h & H_DET)) { h |= H_DET; count_add(cd, RB_A_DETECTED, age, 1); count_add(cd, RB_A_ALL_DETECTED, age, 1); }
            emit_event(G, r, O, c, a, EV_HOSP_CLAIM);
            extra = H_PEND;
        } else {                                     // person_recover, main.pyx:315-318
            h = H_SET_STATE(h, RB_RECOVERED) & ~H_LIST;
            count_add(cd, RB_A_INFECTED, age, -1); count_add(cd, RB_A_RECOVERED, age, 1);
        }
    } else {   // HOSPITALIZED / IN_ICU
        int type;
        if (st == RB_HOSPITALIZED && (sev == RB_CRITICAL || sev == RB_FATAL)) { type = EV_TO_ICU; extra = H_PEND; }   // main.pyx:430-431
        else {
            // person_release_from_hospital, main.pyx:354-367: the outcome does not depend on capacity
            type = st == RB_IN_ICU ? EV_ICU_RELEASE : EV_WARD_RELEASE;
            count_add(cd, st == RB_IN_ICU ? RB_A_IN_ICU : RB_A_IN_WARD, age, -1);
            count_add(cd, RB_A_INFECTED, age, -1);
            if (sev == RB_FATAL) { h = H_SET_STATE(h, RB_DEAD) & ~H_LIST; count_add(cd, RB_A_DEAD, age, 1); count_add(cd, RB_A_NON_HOSPITAL_DEATHS, age, 1); }
            else { h = H_SET_STATE(h, RB_RECOVERED) & ~H_LIST; count_add(cd, RB_A_RECOVERED, age, 1); }
        }
        emit_event(G, r, O, c, a, type);
    }
    lw = h | extra;
    return h;
}

// stage T on one batch of ring T.  Every lane first fetches the authoritative word (flags may have changed behind the
// list's back: detected / queued / vaccinated) and merges the list copy's current day counters into it.
__device__ __forceinline__ void stage_slow(const Eng &G, int r, RepCtr *c, RepCtr *cd, const WarpRings &W, uint32_t head, uint32_t m,
                                           uint32_t next, int lane) {
    const SweepOut O = sweep_out(G, r, c);     // resolved here, not in the caller: the streaming loop stays light on registers
    const size_t base = (size_t)r * G.Npad;
    const bool on = (uint32_t)lane < m;
    bool keep = false, changed = false, removed = false;
    int infected_others = 0;
    uint32_t a = 0, lw = 0, h = 0;
    if (on) {
        a = W.ta[(head + lane) & (SW_RCAP - 1)];
        const uint32_t w = W.tw[(head + lane) & (SW_RCAP - 1)];
        const uint32_t hot = G.hot[base + a];
        if (w & H_PEND) {
            // yesterday's bed / ICU claim was decided by the day boundary, which wrote state and day counter: advance as
            // the reference does on the first sweep after (main.pyx:422-437) -- or the claim failed and the agent is gone
            h = hot & ~H_FRESH;
            if (H_STATE(h) < RB_RECOVERED) { uint32_t dl = H_DL(h); if (dl > 0) dl--; h = H_SET_DL(h, dl); }
        } else h = (hot & ~(H_DAYS_MASK | H_FRESH)) | (w & H_DAYS_MASK);
        const uint32_t st = H_STATE(h);
        if (st >= RB_RECOVERED) {          // R bookkeeping, main.pyx:1969-1972: counted once, on the first sweep after removal
            removed = true;
            infected_others = (int)(G.rec[base + a].cold & 0xffffu);
            G.hot[base + a] = hot | H_INCL;
        } else if (H_DL(h) == 0) {
            h = transition(G, r, c, O, (int32_t)a, h, lw);
            G.hot[base + a] = h;
            keep = true; changed = true;
        } else { lw = h; keep = true; }    // a granted claim with days to go: nothing to write but the list entry
    }
    const uint32_t rm = __ballot_sync(0xffffffffu, removed);
    if (rm) {                              // one pair of atomics per warp batch instead of one per removed agent
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) infected_others += __shfl_xor_sync(0xffffffffu, infected_others, o);
        if (lane == 0) { atomicAdd(&cd->total_infectors, __popc(rm)); if (infected_others) atomicAdd(&cd->total_infections, infected_others); }
    }
    if (keep) list_add(G, r, cd, next, a, lw);      // back of the agent's own segment in tomorrow's buffer (whoever owns its front)
    if (O.upd) {      // sharded mode: the other ranks' copies of this agent learn the new state and flags from the log
        const uint32_t mk = __ballot_sync(0xffffffffu, changed);
        uint32_t b = 0;
        if (lane == 0 && mk) b = atomicAdd(&O.cd->n_upd, (uint32_t)__popc(mk));
        b = __shfl_sync(0xffffffffu, b, 0);
        if (changed) {
            const uint32_t idx = b + __popc(mk & ((1u << lane) - 1u));
            if (idx < O.cap_upd) O.upd[idx] = make_uint2(a, h); else set_problem(O.cd, RB_OTHER_FAILURE);
        }
    }
}

// stage 1 for one list entry: day counters, what is due.  `w` is updated in place; exactly one of keep / want_t is set.
__device__ __forceinline__ void stage_entry(const Eng &G, uint32_t &w, bool &keep, bool &want_e, bool &want_t, uint32_t &desc) {
    const uint32_t st = H_STATE(w);
    if ((w & H_PEND) || st >= RB_RECOVERED) { want_t = true; return; }      // needs the authoritative word / the agent record
    if (w & H_FRESH) { w &= ~H_FRESH; keep = true; return; }                 // infected today before the sweep: wait until tomorrow, main.pyx:402-403
    const uint32_t sev = H_SEV(w), var = H_VAR(w);
    uint32_t dl = H_DL(w);
    if (st == RB_INCUBATION || st == RB_ILLNESS) {
        const int dayidx = st == RB_INCUBATION ? -(int)dl : (int)H_DOI(w);
        if (!(w & H_DET) && dayidx >= -10 && dayidx <= 10 && G.variants[var].iot[dayidx + 10] != 0.0f) {
            want_e = true;
            const uint32_t cls = (st == RB_ILLNESS && sev != RB_ASYMPTOMATIC) ? 1u : 0u;   // factor 0.5, limit 5
            desc = ((uint32_t)(dayidx + 10) << 14) | ((sev == RB_ASYMPTOMATIC ? 1u : 0u) << 19) | (var << 20) | (cls << 22);
        }
        if (st == RB_ILLNESS) { uint32_t doi = H_DOI(w); if (doi < 31) doi++; w = H_SET_DOI(w, doi); }
    }
    if (dl > 0) dl--;
    w = H_SET_DL(w, dl);
    if (dl == 0) want_t = true; else keep = true;
}

// One loop per warp over the segments it owns, 32 entries per step, the next step's entries loaded ahead.  Every stage
// is instantiated exactly ONCE (the heavy stages are thousands of instructions each; a second inlined copy pushes the
// loop out of the instruction cache).
// The sweep of replica r as seen by ONE warp: warp `first` of the `n_warps` that share the replica's segments.
__device__ __forceinline__ void sweep_warp(const Eng &G, const int r, const uint32_t first, const uint32_t n_warps, WarpRings &W, const int lane) {
    RepCtr *c = &G.ctr[r];
    const DevTable *tb = G.tables[c->epoch];
    uint2 *items = G.items + (size_t)r * G.cap_items;
    RepCtr *cd = !G.xbuf ? c : xslot_of(G, G.rank, c->day).hdr;      // counters the sweep adds to
    const uint32_t cur = c->lsel;
    uint32_t e_head = 0, e_tail = 0, t_head = 0, t_tail = 0;

    // One pass per owned segment plus a final pass that only drains the rings: the stages have exactly one call site each.
    uint2 cnt_next = make_uint2(0u, 0u);      // the counters of the next segment are fetched while this one is processed
    if (first < G.n_seg) cnt_next = *seg_count(G, r, cur, first);
    for (uint32_t seg = first; ; seg += n_warps) {
        const bool tail = seg >= G.n_seg;
        const uint2 cnt = cnt_next;
        cnt_next = make_uint2(0u, 0u);
        if (seg + n_warps < G.n_seg) cnt_next = *seg_count(G, r, cur, seg + n_warps);
        const uint32_t nf = min(cnt.x, G.seg_cap), n = min(cnt.x + cnt.y, G.seg_cap);
        const uint2 *in = seg_ptr(G, r, cur, tail ? 0u : seg);
        uint2 *out = seg_ptr(G, r, cur ^ 1u, tail ? 0u : seg);
        uint32_t out_n = 0;
        // entry i of the segment: the front part, then the back part (it ends at the segment's end; its order is irrelevant)
        const uint32_t gap = G.seg_cap - n;
        auto entry = [&](uint32_t i) { return __ldcs(&in[i < nf ? i : i + gap]); };
        uint2 ent = make_uint2(0u, 0u);
        if ((uint32_t)lane < n) ent = entry(lane);
        uint32_t i0 = 0;
        bool again;
        do {
            if (i0 < n) {
                const bool valid = i0 + lane < n;
                const uint2 cur_ent = ent;
                if (i0 + 32 + lane < n) ent = entry(i0 + 32 + lane);      // next step's entry: in flight while this one is processed
                bool keep = false, want_e = false, want_t = false;
                uint32_t w = cur_ent.y, desc = 0;
                if (valid) stage_entry(G, w, keep, want_e, want_t, desc);
                seg_append(G, cd, out, out_n, keep, cur_ent.x, w, lane);
                e_tail = ring_push(W.ea, W.ed, e_tail, want_e, cur_ent.x, desc, lane);
                t_tail = ring_push(W.ta, W.tw, t_tail, want_t, cur_ent.x, w, lane);
                __syncwarp();
            }
            const uint32_t e_av = e_tail - e_head, t_av = t_tail - t_head;
            if (e_av >= 32 || (tail && e_av)) { const uint32_t m = min(32u, e_av); stage_expose(G, r, c, cd, tb, W, e_head, m, items, lane); e_head += m; }
            if (t_av >= 32 || (tail && t_av)) { const uint32_t m = min(32u, t_av); stage_slow(G, r, c, cd, W, t_head, m, cur ^ 1u, lane); t_head += m; }
            __syncwarp();
            i0 += 32;
            again = tail ? (e_tail != e_head || t_tail != t_head) : i0 < n;
        } while (again);
        if (tail) break;
        if (lane == 0) seg_count(G, r, cur ^ 1u, seg)->x = out_n;     // front of tomorrow's segment; its back fills by atomics
    }
}
__global__ void __launch_bounds__(SW_THREADS, SW_CTAS_PER_SM) k_sweep(Eng G) {
    __shared__ WarpRings s_rings[SW_WARPS];
    const int warp = threadIdx.x >> 5;
    sweep_warp(G, blockIdx.y + G.r0, blockIdx.x * SW_WARPS + warp, gridDim.x * SW_WARPS, s_rings[warp], threadIdx.x & 31);
}

// ---------------------------------------------------------------- list maintenance (not on the per-day path)
// k_flush_lists: writes the current day counters of every listed agent back into `hot`, so that `hot` alone describes
// the population (rb_read_agents, rb_save_state).  Harmless at any time between two days: the sweep never reads the day
// counters of `hot` except behind H_PEND, where `hot` is authoritative anyway.
__global__ void k_flush_lists(Eng G) {
    const int r = blockIdx.y;
    const RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const uint32_t cur = c->lsel;
    const int lane = threadIdx.x & 31;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); seg < G.n_seg; seg += n_warps) {
        const uint2 cnt = *seg_count(G, r, cur, seg);
        const uint32_t nf = min(cnt.x, G.seg_cap), n = min(cnt.x + cnt.y, G.seg_cap);
        const uint2 *in = seg_ptr(G, r, cur, seg);
        for (uint32_t i = lane; i < n; i += 32) {
            const uint2 e = in[i < nf ? i : i + (G.seg_cap - n)];
            if ((e.y & H_PEND) || H_STATE(e.y) >= RB_RECOVERED) continue;
            const uint32_t h = G.hot[base + e.x];
            G.hot[base + e.x] = (h & ~(H_DAYS_MASK | H_FRESH)) | (e.y & (H_DAYS_MASK | H_FRESH));
        }
    }
}
// k_rebuild_lists: the active list of every replica from `hot` (after rb_load_state / set_initial_state): everybody who
// is infected, or removed and not yet counted.  A queued agent is detected by the next morning's drain, which runs
// before the next sweep, so its copy carries H_DET already (see transition()).
__global__ void k_rebuild_lists(Eng G) {
    const int r = blockIdx.y;
    RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const uint32_t cur = c->lsel;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < G.N; a += gridDim.x * blockDim.x) {
        if (!owns(G, (uint32_t)a)) continue;
        uint32_t h = G.hot[base + a];
        const uint32_t st = H_STATE(h);
        if (!((st >= RB_INCUBATION && st <= RB_IN_ICU) || (st >= RB_RECOVERED && !(h & H_INCL)))) continue;
        if (h & H_QUEUED) h |= H_DET;
        G.perm[base + a] = sweep_slot(G, c, (uint32_t)a);
        list_add(G, r, c, cur, (uint32_t)a, h);
    }
}
// both buffers' segment counters to zero (rb_create, rb_reset, rb_load_state)
__global__ void k_clear_lists(Eng G) {
    const size_t n = (size_t)G.R * 2 * G.n_seg;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) G.seg_n[i] = make_uint2(0u, 0u);
}

#endif
