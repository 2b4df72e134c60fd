// reina_b200 engine: the per-day agent loop of Reina (cythonsim/main.pyx Context.iterate and everything it
// calls) as hand-written CUDA for sm_100a, behind the C-ABI of include/reina_b200.h.
//
// Data layout (per replica, structure of arrays in HBM, agents stored AGE-SORTED so that age is implied by
// position and a contact target is age_start[band] + u32 % band_size with no indirection):
//   hot[N]      u32  packed word streamed by the daily sweep: state 3b | severity 3b | detected | queued |
//                    variant 2b | fresh | included_in_totals | has_list | vaccinated | days_left 8b | day_of_illness 5b
//   cold[N]     u32  other_people_infected 16b | ward_days 8b | icu_days 8b      (touched by infected agents only)
//   infector[N], first_child[N], next_sib[N], inf_key[N]   infection tree for contact tracing
//   vacc_day[N] i16, winner[N] u64 (atomicMin conflict slots, all-ones when idle)
// Per day (reference order, main.pyx:1994-2016):
//   k_pre     1 CTA / replica   stats row, intervention deltas, imports, test queue drain + contact tracing,
//                               vaccination, sweep start draw
//   k_sweep   grid              Context._iterate_people / person_advance over the packed words; emits contact
//                               work items, capacity events and test-queue entries tagged with sweep position
//   k_expose  grid              one thread per sampled contact: row search, target gather, transmission draw,
//                               atomicMin(winner[target], sweep position of infector | slot)
//   k_resolve grid              winners become infected (severity, incubation draw, infection tree)
//   k_post    1 CTA / replica   beds/ICU first-come-first-served = sort by sweep position + max-plus scan
// Every order-dependent step of the sequential reference is resolved through the agent's sweep position, so the
// result is bit-identical to the sequential CPU oracle and independent of scheduling.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <dlfcn.h>
#include <nccl.h>      // types only: libnccl is loaded at run time by rb_shard_init, single-GPU use never needs it

#include <vector>

#include "../../include/reina_b200.h"
#include "rng.cuh"

#ifndef SW_STREAM_DIV
#define SW_STREAM_DIV 6      // the sweep streams the packed words once more than 1 / 6 of the agents are infected (measured: the bitmap walk wins below)
#endif
#define MAX_INFECTEES 64   // main.pyx:128
#define MAX_CONTACTS 128   // main.pyx:129

// ---------------------------------------------------------------- packed hot word
#define H_STATE(h) ((h) & 7u)
#define H_SEV(h) (((h) >> 3) & 7u)
#define H_DET (1u << 6)
#define H_QUEUED (1u << 7)
#define H_VAR(h) (((h) >> 8) & 3u)
#define H_FRESH (1u << 10)
#define H_INCL (1u << 11)
#define H_LIST (1u << 12)
#define H_VACC (1u << 13)
#define H_DL(h) (((h) >> 14) & 255u)
#define H_DOI(h) (((h) >> 22) & 31u)
#define H_SET_STATE(h, s) (((h) & ~7u) | (uint32_t)(s))
#define H_SET_DL(h, d) (((h) & ~(255u << 14)) | ((uint32_t)(d) << 14))
#define H_SET_DOI(h, d) (((h) & ~(31u << 22)) | ((uint32_t)(d) << 22))

#define KEY_IDLE 0xFFFFFFFFFFFFFFFFull
#define QKEY_SWEEP (1ull << 62)
#define CT_DEAD (1ull << 61)
#define CT_DECIDED (1ull << 60)
#define CT_KEYMASK ((1ull << 60) - 1ull)

enum { EV_HOSP_CLAIM = 0, EV_WARD_RELEASE = 1, EV_TO_ICU = 2, EV_ICU_RELEASE = 3 };

#define SORT_SMEM 2048
#define PRE_THREADS 1024
#define NEG_INF (-(1 << 29))

struct DevTable {
    int32_t n_rows[RB_MAX_AGES];
    float nr_contacts[RB_MAX_AGES];
    double ncdf[RB_MAX_AGES][2][RB_NCDF];
    double cum_p[RB_MAX_AGES][RB_MAX_ROWS];
    uint32_t cum24[RB_MAX_AGES][RB_MAX_ROWS];         // ceil(cum_p * 2^24): for a 24-bit uniform k / 2^24, (k / 2^24 < cum_p) == (k < cum24)
    int32_t start[RB_MAX_AGES][RB_MAX_ROWS];
    int32_t size[RB_MAX_AGES][RB_MAX_ROWS];
    float mask_p[RB_MAX_AGES][RB_MAX_ROWS];
    uint8_t place[RB_MAX_AGES][RB_MAX_ROWS];
    uint8_t lo_age[RB_MAX_AGES][RB_MAX_ROWS], hi_age[RB_MAX_AGES][RB_MAX_ROWS];
    uint8_t susc_uniform[RB_MAX_AGES][RB_MAX_ROWS];   // susceptibility identical for every age of the row's band
    uint8_t guide[RB_MAX_AGES][1024];                 // first row whose cum_p exceeds b/1024: start of the row search
    uint8_t nguide[RB_MAX_AGES][2][256];              // first k with ncdf[k] > b/256: start of the contact-count search
};

struct Attempt { uint32_t cand, parent; unsigned long long key; };

// Everything about one agent that only infections, tracing and capacity outcomes touch, in ONE 32-byte sector:
// an infection then costs one random DRAM sector for the target and one for the infector instead of seven.
struct __align__(32) AgentRec {
    unsigned long long winner;       // atomicMin conflict slot, all-ones when idle
    int32_t infector, first_child, next_sib;   // infection tree (replaces the malloc'd infectees[64], main.pyx:227-233)
    uint32_t inf_key;                // (day << 8) | slot of this agent's infection: orders siblings
    uint32_t cold;                   // other_people_infected 16b | ward_days 8b | icu_days 8b
    int16_t vacc_day, pad;
};

// Per-replica counters.  The scalars the grid kernels hammer with atomics each sit on their own 128-byte line, away
// from the fields every thread only READS (seed, day, sweep start ...): with one big replica all SMs share this one
// struct, and a read that lands on a line with a queue of atomics in front of it waits for all of them.
struct RepCtr {
    int32_t counts[RB_N_ATTRS][RB_MAX_AGES];
    // ---- written by the day-boundary CTA only, read by everybody
    alignas(128) int32_t beds;
    int32_t icu, avail_beds, avail_icu;
    int32_t problem, epoch, testing_mode, day;
    float p_detected_anyway, p_successful_tracing;
    uint32_t seed, start;
    uint32_t fkey[4];
    uint32_t n_queue, qsel;
    uint32_t n_queue_prev;                        // size of the queue drained yesterday (normalises tracing keys for sorting)
    uint32_t any_vacc;                            // set once the first vaccination programme starts
    uint32_t stream_mode;                         // today's sweep streams the packed words instead of the activity bitmap
    uint32_t n_q_base;                            // entries contact tracing put into tomorrow's queue before the sweep
    uint32_t drained;                             // tomorrow's queue was already drained by k_resolve (detections parked in drain_det)
    int32_t ct_cases;
    int32_t vacc_cursor[RB_MAX_VACC];
    // ---- atomics of the grid kernels, one line per group
    alignas(128) uint32_t n_items;
    alignas(128) uint32_t n_succ;
    alignas(128) int32_t exposed_per_day;
    int32_t total_infectors, total_infections;
    alignas(128) uint32_t n_events;
    uint32_t n_newq;
    uint32_t n_upd;                               // population-sharded mode: packed-word updates logged by today's sweep
    alignas(128) int32_t daily_contacts[RB_N_PLACES];
    int32_t by_variant[RB_MAX_VARIANTS];
    alignas(128) int32_t drain_det[RB_MAX_AGES];  // detections of an early queue drain, per age, booked at the next day boundary
    alignas(128) uint32_t n_l0;
    uint32_t n_l1, n_edges;
    long long dbg_t[16];                          // measurement aid: cycles spent per phase of the day-boundary kernel
    long long dbg_last;
};

struct Eng {
    int32_t dbg;                                   // measurement aid: 1/2/3 skip a sweep stage (timing experiments only)
    int32_t N, Npad, n_ages, n_groups, n_variants, R, max_days, row_len, n_import_classes, fhalf;
    uint32_t cap_items, cap_succ, cap_events, cap_queue;
    uint32_t *hot;
    AgentRec *rec;
    uint32_t *sus;                                 // [R][sus_words] 1 bit per agent: still SUSCEPTIBLE (L2-resident gather target)
    int32_t sus_words;
    uint32_t *act;                                 // [R][sus_words] 1 bit per agent: has work in today's sweep (infected, or removed and not yet counted)
    uint2 *items;
    Attempt *succ;
    unsigned long long *ev_key; int32_t *ev_agent;
    unsigned long long *q_key; int32_t *q_agent;   // [R][2][cap_queue]
    RepCtr *ctr;
    int32_t *stats;                                // [R][max_days+1][row_len]
    const rb_day_params *sched;
    DevTable *const *tables;
    const rb_variant *variants;
    const int32_t *age_start;                      // [n_ages+1]
    const uint8_t *age_blk;                        // age of agent (b << 10): coarse index into age_start
    const int32_t *group_of_age;
    const int32_t *import_lo, *import_hi; const float *import_cum;
    // population-sharded mode (rb_shard_init): every rank holds the whole state, sweeps and exposes only the agents
    // it owns, and publishes what the others must know in its slot of the exchange buffer (one all-gather per day)
    int32_t rank, nranks;
    uint8_t *xbuf; size_t xslot;                   // [nranks] message slots; slot `rank` is written locally
    uint32_t xcap_q, xcap_ev, xcap_upd, xcap_succ;
};

// Ownership: stripes of 4096 agents (one warp step of the sweep) dealt round-robin, so every rank holds ~1/nranks of every age.
#define SH_SHIFT 12
__device__ __forceinline__ bool owns(const Eng &G, uint32_t a) { return G.nranks == 1 || (int)((a >> SH_SHIFT) % (uint32_t)G.nranks) == G.rank; }

// One rank's message: a RepCtr used as the header (count deltas of the sweep, list lengths) followed by the lists.
struct XSlot {
    RepCtr *hdr;
    unsigned long long *q_key; int32_t *q_agent;     // test-queue entries created by the sweep
    unsigned long long *ev_key; int32_t *ev_agent;   // capacity events
    uint2 *upd;                                      // (agent, packed word) after a state change
    struct Attempt *succ;                            // successful transmissions
};
__host__ __device__ inline size_t xalign(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ __device__ inline size_t xslot_bytes(uint32_t cq, uint32_t ce, uint32_t cu, uint32_t cs) {
    return xalign(sizeof(RepCtr)) + xalign(8ull * cq) + xalign(4ull * cq) + xalign(8ull * ce) + xalign(4ull * ce) + xalign(8ull * cu) + xalign(16ull * cs);
}
__device__ __forceinline__ XSlot xslot_of(const Eng &G, int rk) {
    uint8_t *p = G.xbuf + (size_t)rk * G.xslot;
    XSlot s;
    s.hdr = (RepCtr *)p; p += xalign(sizeof(RepCtr));
    s.q_key = (unsigned long long *)p; p += xalign(8ull * G.xcap_q);
    s.q_agent = (int32_t *)p; p += xalign(4ull * G.xcap_q);
    s.ev_key = (unsigned long long *)p; p += xalign(8ull * G.xcap_ev);
    s.ev_agent = (int32_t *)p; p += xalign(4ull * G.xcap_ev);
    s.upd = (uint2 *)p; p += xalign(8ull * G.xcap_upd);
    s.succ = (struct Attempt *)p;
    return s;
}

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ int age_of(const Eng &G, int32_t a) {
    // agents are age-sorted: start from the age of the 1024-agent block and walk up (0-1 steps at HUS sizes)
    int age = __ldg(&G.age_blk[a >> 10]);
    while (a >= __ldg(&G.age_start[age + 1])) age++;
    return age;
}
__device__ __forceinline__ int age_in_band(const Eng &G, int32_t a, int lo, int hi) {   // age of agent a, known to lie in [lo, hi]
    hi += 1;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(&G.age_start[mid]) <= a) lo = mid; else hi = mid; }
    return lo;
}
__device__ __forceinline__ uint32_t sweep_pos(const Eng &G, const RepCtr *c, uint32_t a) {
    uint32_t s = feistel(a, (uint32_t)G.N, G.fhalf, c->fkey[0], c->fkey[1], c->fkey[2], c->fkey[3]);
    return s >= c->start ? s - c->start : s + (uint32_t)G.N - c->start;
}
__device__ __forceinline__ void count_add(RepCtr *c, int attr, int age, int d) { atomicAdd(&c->counts[attr][age], d); }
__device__ __forceinline__ void set_problem(RepCtr *c, int p) { atomicCAS(&c->problem, 0, p); }

// Disease.get_symptom_severity, main.pyx:1042-1091 (every FATAL case dies outside hospital, SURVEY 8a note 2)
__device__ __forceinline__ int symptom_severity(const rb_variant *v, int age, float val, bool vacc_eff) {
    float vmod = 1.0f;
    if (vacc_eff) vmod = vmod * 0.1f;
    float syc = v->tab[RB_T_SYMPTOMATIC][age];
    if (val >= syc) return RB_ASYMPTOMATIC;
    syc = syc * vmod;
    float dohc = v->tab[RB_T_DEATH_OUTSIDE_HOSPITAL][age];
    if (dohc != 0.0f) {
        if (val < dohc * syc) return RB_FATAL;
        val = (val - dohc) / (1.0f - dohc);
    }
    float sc = v->tab[RB_T_SEVERE][age], cc = v->tab[RB_T_CRITICAL][age], fc = v->tab[RB_T_FATAL][age];
    if (val < ((fc * cc) * sc) * syc) return RB_FATAL;
    if (val < (cc * sc) * syc) return RB_CRITICAL;
    if (val < sc * syc) return RB_SEVERE;
    return RB_MILD;
}

// person_infect, main.pyx:209-235 + Population.infect :1576-1582.  `src_h` = packed word of the infector
// (ignored when src < 0).  Severity and incubation use wild-type parameters (variant_idx is still 0 there).
__device__ void device_infect(const Eng &G, int r, RepCtr *c, int32_t t, int32_t src, uint32_t src_h, int variant,
                              int slot, bool fresh) {
    const size_t base = (size_t)r * G.Npad;
    const int day = c->day;
    // the two atomics whose results are needed go first; the draws below hide their round trip
    uint32_t old = 0; int32_t prev_child = -1;
    if (src >= 0) {
        old = atomicAdd(&G.rec[base + src].cold, 1u);
        prev_child = atomicExch(&G.rec[base + src].first_child, t);
    }
    const int vd = c->any_vacc ? (int)G.rec[base + t].vacc_day : -1;     // nobody is vaccinated in most configurations
    const int age = age_of(G, t);
    const rb_variant *v0 = &G.variants[0];
    const bool vacc_eff = vd >= 0 && (day - vd) > 14;
    int sev = 0, dl = 0;
    if (owns(G, (uint32_t)t)) {     // severity and day counters are only ever read by the owner's sweep (sharded mode)
        u32x4 x = philox(c->seed, (uint32_t)t, (uint32_t)day, PU_SEVERITY, 0);
        sev = symptom_severity(v0, age, u01f(x.x), vacc_eff);
        dl = clamp255(round_to_int(gamma_f(c->seed, (uint32_t)t, (uint32_t)day, PU_INCUB, v0->incubation_kappa, v0->incubation_theta)));
    }
    if (src >= 0) {
        variant = (int)H_VAR(src_h);
        G.rec[base + t].infector = src;
        if ((src_h & H_LIST) && (old & 0xffffu) >= MAX_INFECTEES) set_problem(c, RB_TOO_MANY_INFECTEES);
        G.rec[base + t].inf_key = ((uint32_t)day << 8) | (uint32_t)slot;
        G.rec[base + t].next_sib = prev_child;
    }
    // a SUSCEPTIBLE agent's word carries nothing but the vaccinated flag, which vacc_day implies
    uint32_t nh = (vd >= 0 ? H_VACC : 0u) | RB_INCUBATION | ((uint32_t)sev << 3) | ((uint32_t)variant << 8) | ((uint32_t)dl << 14);
    if (fresh) nh |= H_FRESH;
    if (c->testing_mode == RB_ALL_WITH_SYMPTOMS_CT) nh |= H_LIST;
    G.hot[base + t] = nh;
    atomicAnd(&G.sus[(size_t)r * G.sus_words + (t >> 5)], ~(1u << (t & 31)));
    atomicOr(&G.act[(size_t)r * G.sus_words + (t >> 5)], 1u << (t & 31));
    count_add(c, RB_A_SUSCEPTIBLE, age, -1);
    count_add(c, RB_A_INFECTED, age, 1);
    count_add(c, RB_A_ALL_INFECTED, age, 1);
    count_add(c, RB_A_NEW_INFECTIONS, age, 1);
    {   // infected_by_variant: one atomic per group of converged lanes with the same variant
        const unsigned act = __activemask();
        const unsigned grp = __match_any_sync(act, variant);
        if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&c->by_variant[variant], __popc(grp));
    }
}

// ---------------------------------------------------------------- block-wide helpers (single CTA)
// Bitonic sort of (key, val) pairs, ascending by key; n <= SORT_SMEM sorts in shared memory, larger lists in
// place in global memory (capacity must be a power of two >= n; the tail is padded with KEY_IDLE).
__device__ void block_sort_pairs(unsigned long long *keys, int32_t *vals, uint32_t n, uint32_t cap,
                                 unsigned long long *sk, int32_t *sv) {
    if (n <= 1) return;
    uint32_t m = 1; while (m < n) m <<= 1;
    if (m <= SORT_SMEM) {
        for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) { sk[i] = i < n ? keys[i] : KEY_IDLE; sv[i] = i < n ? vals[i] : -1; }
        __syncthreads();
        for (uint32_t k = 2; k <= m; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                    uint32_t l = i ^ j;
                    if (l > i) {
                        bool up = (i & k) == 0;
                        unsigned long long a = sk[i], b = sk[l];
                        if ((a > b) == up) { sk[i] = b; sk[l] = a; int32_t t = sv[i]; sv[i] = sv[l]; sv[l] = t; }
                    }
                }
                __syncthreads();
            }
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { keys[i] = sk[i]; vals[i] = sv[i]; }
        __syncthreads();
        return;
    }
    if (m > cap) m = cap;
    for (uint32_t i = n + threadIdx.x; i < m; i += blockDim.x) { keys[i] = KEY_IDLE; vals[i] = -1; }
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                uint32_t l = i ^ j;
                if (l > i) {
                    bool up = (i & k) == 0;
                    unsigned long long a = keys[i], b = keys[l];
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; int32_t t = vals[i]; vals[i] = vals[l]; vals[l] = t; }
                }
            }
            __syncthreads();
        }
}

// inclusive block scan of one int per thread (blockDim.x <= 1024); returns inclusive prefix, *total = block sum
__device__ int block_scan_incl(int v, int *total, int *warp_sums) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_sums[w] = v;
    __syncthreads();
    if (w == 0) {
        int s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        warp_sums[lane] = s;
    }
    __syncthreads();
    int prefix = w > 0 ? warp_sums[w - 1] : 0;
    *total = warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
    return v + prefix;
}

// Order-preserving bucket function for the two kinds of sort keys: capacity events (sweep position << 2 | type) and
// test-queue entries (contact-tracing attempt keys first, then QKEY_SWEEP | sweep position).  Sweep positions are a
// keyed permutation, so B buckets receive ~n/B elements each.
#define SORT_BUCKETS 2048          // bucket counters in shared memory up to here, in global scratch beyond
#define SORT_BUCKETS_MAX 65536
struct BucketMap { uint32_t n_agents, n_prev; int kind; uint32_t B; };     // kind 0: events, 1: queue
__device__ __forceinline__ uint32_t bucket_of(const BucketMap &bm, unsigned long long key) {
    if (bm.kind == 0) return (uint32_t)(((key >> 2) * bm.B) / bm.n_agents);
    if (key & QKEY_SWEEP) return bm.B / 2 + (uint32_t)(((key & 0xffffffffull) * (bm.B / 2)) / bm.n_agents);
    unsigned long long i = key >> 14;                          // queue rank of the tracer, < n_prev
    return (uint32_t)((i * (bm.B / 2)) / (bm.n_prev ? bm.n_prev : 1u));
}

// Ascending sort of n (key, val) pairs with distinct keys by one CTA: counting sort into B ~ n order-preserving
// buckets (histogram + scan), then each bucket (usually 0-2 elements) is put in order by one thread.  O(n) for the
// uniformly spread keys of this engine; `scratch` needs 2 n entries plus B counters.  Returns false if it is too small.
__device__ bool block_bucket_sort(unsigned long long *keys, int32_t *vals, uint32_t n, Attempt *scratch, uint32_t scratch_cap,
                                  BucketMap bm, int32_t *scnt /* shared [SORT_BUCKETS] */, int *warp_sums) {
    uint32_t B = SORT_BUCKETS;
    while (B < n && B < SORT_BUCKETS_MAX) B <<= 1;
    if (2ull * n + B / 4 + 1 > scratch_cap) return false;
    bm.B = B;
    int32_t *cnt = B == SORT_BUCKETS ? scnt : (int32_t *)(scratch + 2ull * n);
    const int tid = threadIdx.x;
    for (uint32_t b = tid; b < B; b += blockDim.x) cnt[b] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += blockDim.x) {
        const unsigned long long k = keys[i];
        const uint32_t b = min(bucket_of(bm, k), B - 1u);
        const uint32_t slot = (uint32_t)atomicAdd(&cnt[b], 1);
        scratch[i].key = k; scratch[i].cand = (uint32_t)vals[i]; scratch[i].parent = b | (slot << 16);
    }
    __syncthreads();
    // exclusive scan of the bucket counts: B / blockDim consecutive buckets per thread
    const uint32_t per = B / blockDim.x, b0 = tid * per;
    int mine = 0;
    for (uint32_t j = 0; j < per; j++) mine += cnt[b0 + j];
    int total;
    int run = block_scan_incl(mine, &total, warp_sums) - mine;
    for (uint32_t j = 0; j < per; j++) { const int v = cnt[b0 + j]; cnt[b0 + j] = run; run += v; }
    __syncthreads();
    Attempt *out = scratch + n;
    for (uint32_t i = tid; i < n; i += blockDim.x) {
        const Attempt e = scratch[i];
        out[cnt[e.parent & 0xffffu] + (e.parent >> 16)] = e;
    }
    __syncthreads();
    for (uint32_t b = tid; b < B; b += blockDim.x) {      // insertion sort inside each bucket
        const uint32_t lo = (uint32_t)cnt[b], hi = b + 1 < B ? (uint32_t)cnt[b + 1] : n;
        for (uint32_t i = lo + 1; i < hi; i++) {
            const Attempt e = out[i];
            uint32_t j = i;
            while (j > lo && out[j - 1].key > e.key) { out[j] = out[j - 1]; j--; }
            out[j] = e;
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += blockDim.x) { keys[i] = out[i].key; vals[i] = (int32_t)out[i].cand; }
    __syncthreads();
    return true;
}

// Context.generate_state, main.pyx:1813-1857: fold per-age counters into age groups + scalars.
__device__ void write_stats_row(const Eng &G, int r, RepCtr *c, int32_t *srow /* shared, >= row_len */) {
    int day = c->day;
    for (int i = threadIdx.x; i < G.row_len; i += blockDim.x) srow[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < RB_N_ATTRS * G.n_ages; i += blockDim.x) {
        int a = i / G.n_ages, age = i - a * G.n_ages;
        int v = c->counts[a][age];
        if (v) atomicAdd(&srow[a * G.n_groups + G.group_of_age[age]], v);
    }
    if (threadIdx.x == 0) {
        int32_t *s = srow + RB_N_ATTRS * G.n_groups;
        s[RB_S_AVAILABLE_ICU] = c->avail_icu; s[RB_S_AVAILABLE_BEDS] = c->avail_beds;
        s[RB_S_TOTAL_ICU] = c->icu; s[RB_S_TOTAL_BEDS] = c->beds;
        s[RB_S_TOTAL_INFECTIONS] = c->total_infections; s[RB_S_TOTAL_INFECTORS] = c->total_infectors;
        s[RB_S_EXPOSED_PER_DAY] = c->exposed_per_day; s[RB_S_CT_CASES_PER_DAY] = c->ct_cases;
        s[RB_S_TABLE_EPOCH] = c->epoch; s[RB_S_DAY] = day;
        for (int i = 0; i < RB_N_PLACES; i++) s[RB_S_CONTACTS0 + i] = c->daily_contacts[i];
        for (int i = 0; i < RB_MAX_VARIANTS; i++) s[RB_S_VARIANT0 + i] = c->by_variant[i];
    }
    __syncthreads();
    int32_t *row = G.stats + ((size_t)r * (G.max_days + 1) + day) * G.row_len;
    for (int i = threadIdx.x; i < G.row_len; i += blockDim.x) row[i] = srow[i];
    __syncthreads();
}

// Population.infect_people + get_import_infection_person, main.pyx:1632-1665.  Sequential semantics: import j takes
// the first of its (up to 10) draws that is SUSCEPTIBLE and was not taken by an earlier import of the same day.
// Up to IMP_CHUNK imports are settled at once, one per thread: each takes its first susceptible draw; if all picks
// of the chunk are distinct (checked through the agents' conflict slots) that IS the sequential result, otherwise
// (probability ~ chunk^2 / N) one thread replays the chunk in order.  Called by the whole CTA.
#define IMP_CHUNK 1024
__device__ __forceinline__ int32_t import_draw(const Eng &G, const RepCtr *c, size_t base, uint32_t ord, uint32_t t) {
    u32x4 x = philox(c->seed, ord, (uint32_t)c->day, PU_IMPORT | (t << 8), 0);
    float p = u01f(x.x);
    int k = G.n_import_classes - 1;
    for (int j = 0; j < G.n_import_classes; j++) if (p <= G.import_cum[j]) { k = j; break; }
    int32_t s = G.age_start[G.import_lo[k]], en = G.age_start[G.import_hi[k] + 1];
    int32_t pi = s + (int32_t)(x.y % (uint32_t)(en - s));
    return H_STATE(G.hot[base + pi]) == RB_SUSCEPTIBLE ? pi : -1;
}
__device__ void import_infections(const Eng &G, int r, RepCtr *c, int count, int variant, int *ordinal, int32_t *chosen /*[IMP_CHUNK]*/,
                                  int *dup_flag) {
    const size_t base = (size_t)r * G.Npad;
    const int tid = threadIdx.x;
    for (int first = 0; first < count; first += IMP_CHUNK) {
        const int m = min(IMP_CHUNK, count - first);
        __syncthreads();
        if (tid == 0) *dup_flag = 0;
        int32_t pick = -1;
        if (tid < m) {
            for (uint32_t t = 0; t < 10 && pick < 0; t++) pick = import_draw(G, c, base, (uint32_t)(*ordinal + first + tid), t);
            chosen[tid] = pick;
            if (pick >= 0) atomicMin(&G.rec[base + pick].winner, (unsigned long long)tid);
        }
        __syncthreads();
        if (pick >= 0 && G.rec[base + pick].winner != (unsigned long long)tid) *dup_flag = 1;
        __syncthreads();
        if (pick >= 0) G.rec[base + pick].winner = KEY_IDLE;
        if (*dup_flag && tid == 0) {
            for (int j = 0; j < m; j++) {
                int32_t found = -1;
                for (uint32_t t = 0; t < 10 && found < 0; t++) {
                    const int32_t pi = import_draw(G, c, base, (uint32_t)(*ordinal + first + j), t);
                    if (pi < 0) continue;
                    bool taken = false;
                    for (int q = 0; q < j; q++) if (chosen[q] == pi) { taken = true; break; }
                    if (!taken) found = pi;
                }
                chosen[j] = found;
            }
        }
        __syncthreads();
        if (tid < m && chosen[tid] >= 0) device_infect(G, r, c, chosen[tid], -1, 0u, variant, 0, true);
    }
    __syncthreads();
    *ordinal += count;
}

// Candidates a tracer reaches (perform_contact_tracing, main.pyx:495-512): slot 0 = its infector, slots 1.. =
// its infectees in infection order (only while the tracer is infected and owns a list, :227-233, :305-307).
__device__ int trace_candidates(const Eng &G, size_t base, int32_t x, uint32_t hx, int32_t *cand /*[65]*/, int *first_slot) {
    int n = 0;
    int32_t inf = G.rec[base + x].infector;
    *first_slot = 1;
    if (inf >= 0) { cand[0] = inf; n = 1; *first_slot = 0; }
    uint32_t st = H_STATE(hx);
    if ((hx & H_LIST) && st >= RB_INCUBATION && st <= RB_IN_ICU) {
        uint32_t keys[MAX_INFECTEES];
        int m = 0;
        int32_t *kids = cand + 1;
        for (int32_t ch = G.rec[base + x].first_child; ch >= 0 && m < MAX_INFECTEES; ch = G.rec[base + ch].next_sib) {
            uint32_t k = G.rec[base + ch].inf_key;
            int j = m++;
            while (j > 0 && keys[j - 1] > k) { keys[j] = keys[j - 1]; kids[j] = kids[j - 1]; j--; }
            keys[j] = k; kids[j] = ch;
        }
        n = 1 + m;
    } else if (inf < 0) n = 0;
    return n;   // valid slots: [*first_slot, n)
}

__device__ __forceinline__ bool trace_eligible(uint32_t h) {   // queue_for_testing guards, main.pyx:476-477
    return H_STATE(h) != RB_DEAD && !(h & (H_DET | H_QUEUED));
}

#define TS(k) do { if (G.dbg == 9 && threadIdx.x == 0) { long long now_ = clock64(); c->dbg_t[k] += now_ - c->dbg_last; c->dbg_last = now_; } } while (0)
// ---------------------------------------------------------------- day boundary (1 CTA per replica)
struct MP { int a, b; };
struct SmemSmall {
    unsigned long long sk[SORT_SMEM];
    int32_t sv[SORT_SMEM];
    union {
        int32_t srow[RB_N_ATTRS * 16 + RB_N_SCALARS];
        struct { MP bed[PRE_THREADS], icu[PRE_THREADS]; } scan;
    } u;
    int warp_sums[32];
    int sh_i[4];
};

__device__ void pre_body(const Eng &G, const int r, SmemSmall &S) {
    unsigned long long *sk = S.sk; int32_t *sv = S.sv; int32_t *srow = S.u.srow; int *warp_sums = S.warp_sums; int *sh_i = S.sh_i;
    RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const int day = c->day;
    const rb_day_params *dp = &G.sched[day];
    const int tid = threadIdx.x;

    if (G.dbg == 9 && threadIdx.x == 0) c->dbg_last = clock64();
    write_stats_row(G, r, c, srow);
    TS(0);

    // apply_intervention effects dated today (main.pyx:1880-1960), then Population.init_day (:1687-1699)
    if (tid == 0) {
        c->testing_mode = dp->testing_mode;
        c->p_detected_anyway = dp->p_detected_anyway;
        c->p_successful_tracing = dp->p_successful_tracing;
        c->beds += dp->beds_delta; c->avail_beds += dp->beds_delta;
        c->icu += dp->icu_delta; c->avail_icu += dp->icu_delta;
    }
    __syncthreads();
    int ordinal = 0;       // uniform across the CTA
    for (int i = 0; i < dp->n_imports; i++)
        import_infections(G, r, c, dp->import_amount[i], dp->import_variant[i], &ordinal, sv, &sh_i[2]);
    __syncthreads();
    for (int i = tid; i < G.n_ages; i += blockDim.x) { c->counts[RB_A_NEW_INFECTIONS][i] = 0; c->counts[RB_A_DETECTED][i] = 0; }
    if (tid < RB_N_PLACES) c->daily_contacts[tid] = 0;
    if (tid < RB_MAX_VARIANTS) c->by_variant[tid] = 0;
    __syncthreads();
    if (tid == 0) { c->epoch = dp->table_epoch; c->total_infectors = 0; c->total_infections = 0; c->exposed_per_day = 0; }
    for (int v = 0; v < G.n_variants; v++)
        if (dp->trickle[v]) import_infections(G, r, c, dp->trickle[v], v, &ordinal, sv, &sh_i[2]);
    __syncthreads();

    TS(1);   // imports + init_day
    // HealthcareSystem.iterate (main.pyx:514-558): drain yesterday's queue
    const uint32_t cur = c->qsel, nxt = cur ^ 1u;
    unsigned long long *qk = G.q_key + ((size_t)r * 2 + cur) * G.cap_queue;
    int32_t *qa = G.q_agent + ((size_t)r * 2 + cur) * G.cap_queue;
    unsigned long long *nk = G.q_key + ((size_t)r * 2 + nxt) * G.cap_queue;
    int32_t *na = G.q_agent + ((size_t)r * 2 + nxt) * G.cap_queue;
    const uint32_t nq = c->n_queue;
    const bool ct = c->testing_mode == RB_ALL_WITH_SYMPTOMS_CT;
    if (tid == 0) { c->ct_cases = (int32_t)nq; c->n_newq = 0; c->n_l0 = 0; c->n_l1 = 0; c->n_edges = 0; }
    if (ct && nq > 1) {                                            // queue order only matters for tracing
        BucketMap bm; bm.n_agents = (uint32_t)G.N; bm.n_prev = c->n_queue_prev; bm.kind = 1;
        if (!block_bucket_sort(qk, qa, nq, G.succ + (size_t)r * G.cap_succ, G.cap_succ, bm, sv, warp_sums))
            block_sort_pairs(qk, qa, nq, G.cap_queue, sk, sv);
    }
    __syncthreads();
    if (c->drained) {      // k_resolve already marked the queued agents detected: book the counts here, where the reference drains
        for (int age = tid; age < G.n_ages; age += blockDim.x) {
            const int d = c->drain_det[age];
            if (d) { c->counts[RB_A_DETECTED][age] += d; c->counts[RB_A_ALL_DETECTED][age] += d; c->drain_det[age] = 0; }
        }
    } else
    for (uint32_t i = tid; i < nq; i += blockDim.x) {
        int32_t a = qa[i];
        uint32_t h = G.hot[base + a];
        if (h & H_DET) set_problem(c, RB_WRONG_STATE);   // person_detect, main.pyx:294-298
        G.hot[base + a] = (h & ~H_QUEUED) | H_DET;
        int age = age_of(G, a);
        count_add(c, RB_A_DETECTED, age, 1); count_add(c, RB_A_ALL_DETECTED, age, 1);
    }
    __syncthreads();
    if (tid == 0) c->drained = 0u;

    TS(2);   // queue drain
    if (ct && nq > 0) {
        // Depth-first contact tracing resolved in parallel.  Attempt key = (queue rank, level-0 slot, level-1 slot);
        // an attempt queues its candidate iff it is the smallest-key LIVE attempt on it that EXISTS; a level-1
        // attempt exists iff its tracer was itself queued by a level-0 attempt (main.pyx:498-499, 505-512).
        Attempt *l0 = G.succ + (size_t)r * G.cap_succ;
        Attempt *l1 = (Attempt *)(G.items + (size_t)r * G.cap_items);
        const uint32_t cap_l1 = G.cap_items / 2;
        const float ptr = c->p_successful_tracing;
        for (uint32_t i = tid; i < nq; i += blockDim.x) {
            int32_t x = qa[i];
            int32_t cand[MAX_INFECTEES + 1]; int first;
            int n = trace_candidates(G, base, x, G.hot[base + x], cand, &first);
            for (int a = first; a < n; a++) {
                int32_t cc = cand[a];
                if (!trace_eligible(G.hot[base + cc])) continue;
                u32x4 rx = philox(c->seed, (uint32_t)x, (uint32_t)day, PU_TRACE, (uint32_t)cc);
                if (!chance(u01d(rx.x, rx.y), ptr)) continue;
                unsigned long long key = ((unsigned long long)i << 14) | ((unsigned long long)a << 7);
                uint32_t idx = atomicAdd(&c->n_l0, 1u);
                if (idx < G.cap_succ) { l0[idx].cand = (uint32_t)cc; l0[idx].parent = (uint32_t)x; l0[idx].key = key; atomicMin(&G.rec[base + (cc)].winner, key); }
                else set_problem(c, RB_OTHER_FAILURE);
            }
        }
        __syncthreads();
        uint32_t n0 = min(c->n_l0, G.cap_succ);
        for (uint32_t j = tid; j < n0; j += blockDim.x) {
            Attempt at = l0[j];
            if (G.rec[base + (at.cand)].winner != at.key) continue;
            int32_t x = (int32_t)at.cand;
            int32_t cand[MAX_INFECTEES + 1]; int first;
            int n = trace_candidates(G, base, x, G.hot[base + x], cand, &first);
            for (int b = first; b < n; b++) {
                int32_t cc = cand[b];
                if (!trace_eligible(G.hot[base + cc])) continue;
                u32x4 rx = philox(c->seed, (uint32_t)x, (uint32_t)day, PU_TRACE, (uint32_t)cc);
                if (!chance(u01d(rx.x, rx.y), ptr)) continue;
                uint32_t idx = atomicAdd(&c->n_l1, 1u);
                if (idx < cap_l1) { l1[idx].cand = (uint32_t)cc; l1[idx].parent = (uint32_t)x; l1[idx].key = at.key | (unsigned long long)(b + 1); }
                else set_problem(c, RB_OTHER_FAILURE);
            }
        }
        __syncthreads();
        uint32_t n1 = min(c->n_l1, cap_l1);
        // kill edges: a level-1 attempt that precedes the level-0 winner of the same candidate
        uint32_t *esrc = (uint32_t *)(G.ev_key + (size_t)r * G.cap_events);
        uint32_t *edst = (uint32_t *)(G.ev_agent + (size_t)r * G.cap_events);
        const uint32_t cap_e = G.cap_events;
        for (uint32_t k = tid; k < n1; k += blockDim.x) {
            unsigned long long w = G.rec[base + (l1[k].cand)].winner;
            if (w != KEY_IDLE && l1[k].key < w) {
                uint32_t idx = atomicAdd(&c->n_edges, 1u);
                if (idx < cap_e) { esrc[idx] = l1[k].parent; edst[idx] = l1[k].cand; } else set_problem(c, RB_OTHER_FAILURE);
            }
        }
        __syncthreads();
        uint32_t ne = min(c->n_edges, cap_e);
        if (ne > 0 && tid == 0) {
            // decide candidates in increasing order of their level-0 key: a candidate loses its tracing rights iff
            // some level-1 attempt from a tracer that kept its rights precedes its own level-0 attempt
            for (;;) {
                unsigned long long best = KEY_IDLE; uint32_t bd = 0;
                for (uint32_t k = 0; k < ne; k++) {
                    unsigned long long w = G.rec[base + (edst[k])].winner;
                    if (!(w & CT_DECIDED) && (w & CT_KEYMASK) < best) { best = w & CT_KEYMASK; bd = edst[k]; }
                }
                if (best == KEY_IDLE) break;
                bool dead = false;
                for (uint32_t k = 0; k < ne; k++) if (edst[k] == bd && !(G.rec[base + (esrc[k])].winner & CT_DEAD)) { dead = true; break; }
                G.rec[base + (bd)].winner = best | CT_DECIDED | (dead ? CT_DEAD : 0ull);
            }
            for (uint32_t k = 0; k < ne; k++) {
                unsigned long long w = G.rec[base + (edst[k])].winner;
                if (w == KEY_IDLE) continue;
                G.rec[base + (edst[k])].winner = (w & CT_DEAD) ? KEY_IDLE : (w & CT_KEYMASK);
            }
        }
        __syncthreads();
        for (uint32_t k = tid; k < n1; k += blockDim.x) {
            Attempt e = l1[k];
            if (G.rec[base + (e.parent)].winner == (e.key & ~127ull)) atomicMin(&G.rec[base + (e.cand)].winner, e.key);
        }
        __syncthreads();
        for (uint32_t j = tid; j < n0 + n1; j += blockDim.x) {
            Attempt e = j < n0 ? l0[j] : l1[j - n0];
            if (G.rec[base + (e.cand)].winner != e.key) continue;
            if (j >= n0 && G.rec[base + (e.parent)].winner != (e.key & ~127ull)) continue;
            uint32_t idx = atomicAdd(&c->n_newq, 1u);
            if (idx < G.cap_queue) { nk[idx] = e.key; na[idx] = (int32_t)e.cand; } else set_problem(c, RB_OTHER_FAILURE);
            G.hot[base + e.cand] |= H_QUEUED;
        }
        __syncthreads();
        for (uint32_t j = tid; j < n0 + n1; j += blockDim.x) { Attempt e = j < n0 ? l0[j] : l1[j - n0]; G.rec[base + (e.cand)].winner = KEY_IDLE; }
        __syncthreads();
    }

    TS(3);   // contact tracing
    // vaccinate_people (main.pyx:560-583): top-down walk of the age-sorted range; eligibility only ever turns
    // off (dead / vaccinated / detected), so a per-programme cursor below which the walk resumes is exact.
    for (int p = 0; p < dp->n_vacc; p++) {
        int nr = dp->vacc_nr[p];
        if (!nr) continue;
        if (tid == 0) c->any_vacc = 1u;
        int slot = dp->vacc_slot[p];
        int32_t s = G.age_start[dp->vacc_min_age[p]], en = G.age_start[dp->vacc_max_age[p] + 1];
        if (nr > en - s) nr = en - s;
        int32_t pos = c->vacc_cursor[slot];                 // -2 = programme not started yet
        if (pos == -2 || pos > en - 1) pos = en - 1;
        int done = 0;
        __syncthreads();
        while (done < nr && pos >= s) {
            int32_t idx = pos - tid;
            uint32_t h = 0; bool el = false;
            if (idx >= s) { h = G.hot[base + idx]; el = H_STATE(h) != RB_DEAD && !(h & (H_VACC | H_DET)); }
            int total;
            int rank = block_scan_incl(el ? 1 : 0, &total, warp_sums);
            int want = nr - done;
            if (el && rank <= want) {
                G.hot[base + idx] = h | H_VACC;
                G.rec[base + idx].vacc_day = (int16_t)day;
                count_add(c, RB_A_VACCINATED, age_of(G, idx), 1);
                if (rank == want) sh_i[1] = idx - 1;      // walk stops right below the last person vaccinated
            }
            __syncthreads();
            if (total >= want) { done = nr; pos = sh_i[1]; }
            else { done += total; pos -= (int32_t)blockDim.x; }
            __syncthreads();
        }
        if (tid == 0) c->vacc_cursor[slot] = pos < s - 1 ? s - 1 : pos;
    }
    __syncthreads();

    TS(4);   // vaccination
    if (tid == 0) {
        u32x4 x = philox(c->seed, 0u, (uint32_t)day, PU_START, 0);     // _iterate_people, main.pyx:1988
        c->start = x.x % (uint32_t)G.N;
        c->n_items = 0; c->n_succ = 0; c->n_events = 0;
        c->n_queue_prev = nq;      // tomorrow's queue holds tracing keys whose rank field is < nq
        // dense days (> 1/24 of the agents infected) stream the packed words, sparse days walk the activity bitmap
        int infected = 0;
        for (int age = 0; age < G.n_ages; age++) infected += c->counts[RB_A_INFECTED][age];
        c->stream_mode = (long long)infected * SW_STREAM_DIV > (long long)G.N ? 1u : 0u;
        c->n_q_base = c->n_newq;
    }
    if (G.xbuf) {     // this rank's message header: the sweep and the contact kernel add to it from zero
        uint32_t *hw = (uint32_t *)xslot_of(G, G.rank).hdr;
        for (int i = tid; i < (int)(sizeof(RepCtr) / 4); i += blockDim.x) hw[i] = 0u;
    }
}

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
#define SW_CTAS_PER_SM 8
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
        const XSlot x = xslot_of(G, G.rank);
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
    const int r = blockIdx.y;
    RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const DevTable *tb = G.tables[c->epoch];
    uint2 *items = G.items + (size_t)r * G.cap_items;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpRings &W = s_rings[warp];
    RepCtr *cd = !G.xbuf ? c : xslot_of(G, G.rank).hdr;      // counters the sweep adds to
    const int nrk = G.nranks, rk = G.rank;
    const int stride = gridDim.x * SW_WARPS;
    const bool stream = c->stream_mode != 0;
    uint32_t head = 0, tail = 0, e_head = 0, e_tail = 0, t_head = 0, t_tail = 0;

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
        while (more && tail - head < 32) {
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
                    wb = make_uint4(0, 0, 0, 0);
                    if (vi < n_vec) wb = __ldg(&act4[vi]);
                    a0 = (uint32_t)vi * 128u;
                    if (__any_sync(0xffffffffu, (wb.x | wb.y | wb.z | wb.w) != 0u)) { part = 0; cw = wb.x; }
                    else more = j < n_mine;
                } else if (!__any_sync(0xffffffffu, cw != 0u)) {
                    part++;
                    cw = part == 1 ? wb.y : (part == 2 ? wb.z : wb.w);
                    if (part == 4) more = j < n_mine;
                } else {
                    // every lane queues up to 4 of its set bits per round: at most 128 pushes, the ring holds 256
                    const uint32_t mine = min(__popc(cw), 4);
                    uint32_t incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                    uint32_t p = tail + incl - mine;
                    // the active agents' packed words are gathered HERE (the only per-agent gather of the sweep): up to four
                    // independent loads per lane in flight instead of one per lane in the consumer
                    uint32_t ia[4], wa[4];
#pragma unroll
                    for (uint32_t k = 0; k < 4; k++)
                        if (k < mine) { ia[k] = a0 + (uint32_t)part * 32u + (uint32_t)(__ffs(cw) - 1); cw &= cw - 1u; wa[k] = G.hot[base + ia[k]]; }
#pragma unroll
                    for (uint32_t k = 0; k < 4; k++)
                        if (k < mine) { W.qi[(p + k) & (SW_QCAP - 1)] = ia[k]; W.qw[(p + k) & (SW_QCAP - 1)] = wa[k]; }
                    tail += tot;
                }
            }
            __syncwarp();
        }
        // ---------------- consume: full batches while the source lasts, whatever is left afterwards
        const uint32_t av = tail - head;
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

// ---------------------------------------------------------------- k_expose
// Contacts.  One thread per group of four contact slots of one infector (they share one Philox block):
// get_one_contact (main.pyx:1290-1304) picks the row, daily_contacts[place] is counted, and a coarse 8-bit filter
// (thinning, see the oracle) decides whether the contact can transmit at all.  The few survivors go through a
// warp-private shared-memory ring and are finished on dense warps: get_person_from_age_range (:1525-1535),
// person_expose / did_infect (:238-244, 908-934), atomicMin on the target's conflict slot.
#ifndef EX_THREADS
#define EX_THREADS 128
#endif
#ifndef EX_CTAS_PER_SM
#define EX_CTAS_PER_SM 12        // CTAs per SM the grid is sized for: all resident (40 registers), one wave
#endif
#define EX_WARPS (EX_THREADS / 32)
#define EX_RCAP 256

__device__ __forceinline__ void expose_survivors(const Eng &G, int r, RepCtr *c, RepCtr *cd, const DevTable *tb, const uint2 *items, const uint32_t *sus,
                                                 Attempt *succ, uint32_t cap_succ, const uint32_t *ri, const uint32_t *rx, uint32_t head, uint32_t m, int lane) {
    const size_t base = (size_t)r * G.Npad;
    bool ok = false;
    uint32_t a = 0, t = 0, slot = 0;
    if ((uint32_t)lane < m) {
        const uint2 it = items[ri[(head + lane) & (EX_RCAP - 1)]];
        const uint32_t info = rx[(head + lane) & (EX_RCAP - 1)];
        const uint32_t row = (info >> 7) & 127u, kq = info >> 14;
        slot = info & 127u;
        a = it.x;
        const uint32_t age = (it.y >> 7) & 127u, dayidx = (it.y >> 14) & 31u, var = (it.y >> 20) & 3u;
        const rb_variant *v = &G.variants[var];
        float si = v->iot[dayidx];
        if ((it.y >> 19) & 1u) si = si * v->p_asymptomatic_infection;
        const u32x4 y = philox(c->seed, a, (uint32_t)c->day, PU_CONTACT2 | (slot << 8), 0);
        t = (uint32_t)tb->start[age][row] + y.x % (uint32_t)tb->size[age][row];
        // person_expose (main.pyx:238-244): only a SUSCEPTIBLE target can be infected; the 1-bit-per-agent map keeps
        // this random gather inside L2 instead of pulling a 32-byte DRAM sector per contact
        if ((__ldg(&sus[t >> 5]) >> (t & 31)) & 1u) {
            const int tage = tb->susc_uniform[age][row] ? (int)tb->lo_age[age][row]
                                                        : age_in_band(G, (int32_t)t, tb->lo_age[age][row], tb->hi_age[age][row]);
            const float pr = (si * v->tab[RB_T_SUSCEPTIBILITY][tage]) * v->infectiousness_multiplier;
            if (((double)y.y * (1.0 / 4294967296.0)) * (double)kq < (double)pr * 256.0) {
                ok = true;
                const float mp = tb->mask_p[age][row];
                if (mp != 0.0f) {
                    const float ma = mp * v->p_mask_protects_others, mb = mp * v->p_mask_protects_wearer;
                    const float pm = (ma + mb) - ma * mb;
                    if (chance((double)y.z * (1.0 / 4294967296.0), pm)) ok = false;
                }
            }
        }
    }
    const uint32_t okm = __ballot_sync(0xffffffffu, ok);
    if (!okm) return;
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(&cd->n_succ, (uint32_t)__popc(okm));     // one atomic per warp batch
    b = __shfl_sync(0xffffffffu, b, 0);
    if (!ok) return;
    const uint32_t idx = b + __popc(okm & ((1u << lane) - 1u));
    const unsigned long long key = ((unsigned long long)sweep_pos(G, c, a) << 7) | slot;
    if (idx < cap_succ) {
        succ[idx].cand = t; succ[idx].parent = a; succ[idx].key = key;
        if (!G.xbuf) atomicMin(&G.rec[base + t].winner, key);     // sharded: k_merge does it over every rank's list
    } else set_problem(cd, RB_OTHER_FAILURE);
}

__global__ void __launch_bounds__(EX_THREADS) k_expose(Eng G) {
    __shared__ int s_place[RB_N_PLACES];
    __shared__ uint32_t s_ri[EX_WARPS][EX_RCAP], s_rx[EX_WARPS][EX_RCAP];
    const int r = blockIdx.y;
    RepCtr *c = &G.ctr[r];
    const DevTable *tb = G.tables[c->epoch];
    const uint32_t n = min(c->n_items, G.cap_items);
    if (blockIdx.x * blockDim.x >= n) return;
    if (threadIdx.x < RB_N_PLACES) s_place[threadIdx.x] = 0;
    __syncthreads();
    const uint2 *items = G.items + (size_t)r * G.cap_items;
    Attempt *succ = G.succ + (size_t)r * G.cap_succ;
    uint32_t cap_succ = G.cap_succ;
    RepCtr *cd = c;
    if (G.xbuf) { const XSlot x = xslot_of(G, G.rank); succ = x.succ; cap_succ = G.xcap_succ; cd = x.hdr; }
    const uint32_t *sus = G.sus + (size_t)r * G.sus_words;
    const uint32_t day = (uint32_t)c->day;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *ri = s_ri[warp], *rx = s_rx[warp];
    uint32_t head = 0, tail = 0;
    uint32_t places = 0;                 // this thread's per-place counters, 5 bits each, flushed every 7 iterations
    int since_flush = 0;
    for (uint32_t i0 = blockIdx.x * blockDim.x + warp * 32; i0 < n; i0 += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + lane;
        const bool valid = i < n;
        uint32_t words[4] = {0, 0, 0, 0}, ncnt = 0, age = 0, grp = 0;
        int kq = 0;
        if (valid) {
            const uint2 it = items[i];
            grp = it.y & 31u; ncnt = ((it.y >> 5) & 3u) + 1u; age = (it.y >> 7) & 127u;
            const rb_variant *v = &G.variants[(it.y >> 20) & 3u];
            float si = v->iot[(it.y >> 14) & 31u];
            if ((it.y >> 19) & 1u) si = si * v->p_asymptomatic_infection;
            const float p_upper = (si * v->reserved[0]) * v->infectiousness_multiplier;
            kq = (int)(p_upper * 256.0f) + 1;
            if (kq > 256) kq = 256;
            const u32x4 x = philox(c->seed, it.x, day, PU_CONTACT | (grp << 8), 0);
            words[0] = x.x; words[1] = x.y; words[2] = x.z; words[3] = x.w;
        }
        const int nrows = tb->n_rows[age];
        const uint32_t *cum24 = tb->cum24[age];
#pragma unroll
        for (uint32_t w = 0; w < 4; w++) {
            bool pass = false;
            uint32_t row = 0;
            if (w < ncnt) {
                const uint32_t word = words[w];
                // u = (word >> 8) / 2^24; linear scan for the first row with u < cum_p, as an integer compare against
                // ceil(cum_p 2^24); rows below guide[u's top 10 bits] cannot match
                const uint32_t k24 = word >> 8;
                row = tb->guide[age][word >> 22];
                while ((int)row < nrows - 1 && !(k24 < cum24[row])) row++;   // last row on overrun: the reference fails there (p ~ 1e-15)
                places += 1u << (5 * tb->place[age][row]);               // daily_contacts[place]++ (main.pyx:1571)
                pass = (int)(word & 255u) < kq;
            }
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                const uint32_t p = (tail + __popc(m & ((1u << lane) - 1u))) & (EX_RCAP - 1);
                ri[p] = i; rx[p] = (grp * 4u + w) | (row << 7) | ((uint32_t)kq << 14);
            }
            tail += __popc(m);
        }
        __syncwarp();
        while (tail - head >= 32) { expose_survivors(G, r, c, cd, tb, items, sus, succ, cap_succ, ri, rx, head, 32, lane); head += 32; }
        __syncwarp();
        if (++since_flush == 7) {        // 7 iterations x 4 contacts = 28 < 32 fits the 5-bit fields
#pragma unroll
            for (int pl = 0; pl < RB_N_PLACES; pl++) { const uint32_t k = (places >> (5 * pl)) & 31u; if (k) atomicAdd(&s_place[pl], (int)k); }
            places = 0; since_flush = 0;
        }
    }
    if (tail != head) expose_survivors(G, r, c, cd, tb, items, sus, succ, cap_succ, ri, rx, head, tail - head, lane);
#pragma unroll
    for (int pl = 0; pl < RB_N_PLACES; pl++) { const uint32_t k = (places >> (5 * pl)) & 31u; if (k) atomicAdd(&s_place[pl], (int)k); }
    __syncthreads();
    if (threadIdx.x < RB_N_PLACES && s_place[threadIdx.x]) atomicAdd(&cd->daily_contacts[threadIdx.x], s_place[threadIdx.x]);
}

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

// ---------------------------------------------------------------- k_resolve
// DRAIN: this day is followed by the fused day boundary, so tomorrow's test queue -- complete once today's sweep is
// over -- is drained here by the whole grid instead of by tomorrow's single boundary CTA (HealthcareSystem.iterate,
// main.pyx:514-545: every queued agent is detected).  The per-age detection counts are parked in drain_det and booked
// by the boundary at the point where the reference drains, so every stats row is unchanged.
template <bool DRAIN>
__global__ void __launch_bounds__(256) k_resolve(Eng G) {
    const int r = blockIdx.y;
    RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const uint32_t n = min(c->n_succ, G.cap_succ);
    const Attempt *succ = G.succ + (size_t)r * G.cap_succ;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Attempt at = succ[i];
        const unsigned long long w = G.rec[base + at.cand].winner;
        const uint32_t src_h = G.hot[base + at.parent];       // in flight together with the conflict slot
        if (w != at.key) continue;                             // first infector in sweep order wins
        device_infect(G, r, c, (int32_t)at.cand, (int32_t)at.parent, src_h, 0, (int)(at.key & 127ull), false);
        G.rec[base + at.cand].winner = KEY_IDLE;
    }
    if (DRAIN) {
        const uint32_t nq = min(c->n_newq, G.cap_queue);
        const int32_t *qa = G.q_agent + ((size_t)r * 2 + (c->qsel ^ 1u)) * G.cap_queue;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
            const int32_t a = qa[i];
            const uint32_t h = G.hot[base + a];
            if (h & H_DET) set_problem(c, RB_WRONG_STATE);   // person_detect, main.pyx:294-298
            G.hot[base + a] = (h & ~H_QUEUED) | H_DET;
            atomicAdd(&c->drain_det[age_of(G, a)], 1);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) c->drained = 1u;
    }
}

// ---------------------------------------------------------------- k_post
// HealthcareSystem.hospitalize / release / to_icu / release_from_icu (main.pyx:617-651) are first-come-first-served
// in sweep order.  Each event is a map x -> max(x + a, b) on the free-bed (and free-ICU) counter; sorting the day's
// events by sweep position and scanning the composed maps gives every claim the counter value it would have seen.
__device__ __forceinline__ MP mp_compose(MP f, MP g) {   // apply f, then g
    MP o; o.a = f.a + g.a; int t = f.b + g.a; o.b = t > g.b ? t : g.b; if (o.b < NEG_INF) o.b = NEG_INF; return o;
}
__device__ __forceinline__ MP mp_bed(int type) {
    MP m; m.a = 0; m.b = NEG_INF;
    if (type == EV_HOSP_CLAIM) { m.a = -1; m.b = 0; } else if (type == EV_WARD_RELEASE || type == EV_TO_ICU) m.a = 1;
    return m;
}
__device__ __forceinline__ MP mp_icu(int type) {
    MP m; m.a = 0; m.b = NEG_INF;
    if (type == EV_TO_ICU) { m.a = -1; m.b = 0; } else if (type == EV_ICU_RELEASE) m.a = 1;
    return m;
}
__device__ __forceinline__ int mp_apply(MP f, int x) { int t = x + f.a; return t > f.b ? t : f.b; }

__device__ void post_body(const Eng &G, const int r, SmemSmall &S) {
    unsigned long long *sk = S.sk; int32_t *sv = S.sv; MP *s_bed = S.u.scan.bed, *s_icu = S.u.scan.icu;
    RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const int tid = threadIdx.x;
    const uint32_t n = min(c->n_events, G.cap_events);
    unsigned long long *ek = G.ev_key + (size_t)r * G.cap_events;
    int32_t *ea = G.ev_agent + (size_t)r * G.cap_events;
    const int day = c->day;
    if (G.dbg == 9 && threadIdx.x == 0) c->dbg_last = clock64();
    if (n > 0) {
        BucketMap bm; bm.n_agents = (uint32_t)G.N; bm.n_prev = 0; bm.kind = 0;
        if (n > 1 && !block_bucket_sort(ek, ea, n, G.succ + (size_t)r * G.cap_succ, G.cap_succ, bm, sv, S.warp_sums))
            block_sort_pairs(ek, ea, n, G.cap_events, sk, sv);
        __syncthreads();
        TS(8);   // event sort
        const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
        const uint32_t lo = min(n, tid * per), hi = min(n, lo + per);
        MP fb; fb.a = 0; fb.b = NEG_INF; MP fi = fb;
        for (uint32_t i = lo; i < hi; i++) { int type = (int)(ek[i] & 3ull); fb = mp_compose(fb, mp_bed(type)); fi = mp_compose(fi, mp_icu(type)); }
        s_bed[tid] = fb; s_icu[tid] = fi;
        __syncthreads();
        for (int o = 1; o < (int)blockDim.x; o <<= 1) {     // Hillis-Steele inclusive scan of composed maps
            MP pb, pi; bool has = tid >= o;
            if (has) { pb = s_bed[tid - o]; pi = s_icu[tid - o]; }
            __syncthreads();
            if (has) { s_bed[tid] = mp_compose(pb, s_bed[tid]); s_icu[tid] = mp_compose(pi, s_icu[tid]); }
            __syncthreads();
        }
        const int beds0 = c->avail_beds, icu0 = c->avail_icu;
        int beds = tid > 0 ? mp_apply(s_bed[tid - 1], beds0) : beds0;
        int icu = tid > 0 ? mp_apply(s_icu[tid - 1], icu0) : icu0;
        for (uint32_t i = lo; i < hi; i++) {
            int type = (int)(ek[i] & 3ull);
            int32_t a = ea[i];
            if (type == EV_HOSP_CLAIM || type == EV_TO_ICU) {
                uint32_t h = G.hot[base + a];
                const uint32_t sev = H_SEV(h);
                const rb_variant *v = &G.variants[H_VAR(h)];
                const int age = age_of(G, a);
                const uint32_t cold = G.rec[base + a].cold;
                const bool ok = type == EV_HOSP_CLAIM ? beds > 0 : icu > 0;
                bool dies = false;
                if (!ok) {      // Disease.dies_in_hospital(care_available=False), main.pyx:957-974
                    if (sev == RB_FATAL) dies = true;
                    else {
                        float ch = sev == RB_CRITICAL ? v->p_icu_death_no_beds : (sev == RB_SEVERE ? v->p_hospital_death_no_beds : 0.0f);
                        u32x4 x = philox(c->seed, (uint32_t)a, (uint32_t)day, PU_NOBED, 0);
                        dies = chance(u01d(x.x, x.y), ch);
                    }
                }
                if (type == EV_HOSP_CLAIM) {            // person_hospitalize, main.pyx:327-338
                    if (ok) { h = H_SET_DL(H_SET_STATE(h, RB_HOSPITALIZED), (cold >> 16) & 255u); count_add(c, RB_A_IN_WARD, age, 1); }
                    else {
                        count_add(c, RB_A_INFECTED, age, -1);
                        if (dies) { h = H_SET_STATE(h, RB_DEAD) & ~H_LIST; count_add(c, RB_A_DEAD, age, 1); if (sev == RB_FATAL) count_add(c, RB_A_NON_HOSPITAL_DEATHS, age, 1); }
                        else { h = H_SET_STATE(h, RB_RECOVERED) & ~H_LIST; count_add(c, RB_A_RECOVERED, age, 1); }
                    }
                } else {                                 // person_transfer_to_icu, main.pyx:341-351
                    count_add(c, RB_A_IN_WARD, age, -1);
                    if (!ok && dies) {
                        count_add(c, RB_A_INFECTED, age, -1); count_add(c, RB_A_DEAD, age, 1);
                        if (sev == RB_FATAL) count_add(c, RB_A_NON_HOSPITAL_DEATHS, age, 1);
                        h = H_SET_STATE(h, RB_DEAD) & ~H_LIST;
                    } else {
                        h = H_SET_DL(H_SET_STATE(h, RB_IN_ICU), (cold >> 24) & 255u);
                        count_add(c, RB_A_IN_ICU, age, 1); count_add(c, RB_A_CUM_ICU, age, 1);
                    }
                }
                G.hot[base + a] = h;
            }
            beds = mp_apply(mp_bed(type), beds);
            icu = mp_apply(mp_icu(type), icu);
        }
        __syncthreads();
        if (tid == 0) { c->avail_beds = mp_apply(s_bed[blockDim.x - 1], beds0); c->avail_icu = mp_apply(s_icu[blockDim.x - 1], icu0); }
    }
    __syncthreads();
    TS(9);   // capacity scan + outcomes
    if (tid == 0) {
        c->qsel ^= 1u;
        c->n_queue = min(c->n_newq, G.cap_queue);
        c->n_newq = 0;
        c->day = day + 1;            // main.pyx:2009
    }
}

__global__ void __launch_bounds__(PRE_THREADS) k_pre(Eng G) { __shared__ SmemSmall S; pre_body(G, blockIdx.x, S); }
__global__ void __launch_bounds__(PRE_THREADS) k_post(Eng G) { __shared__ SmemSmall S; post_body(G, blockIdx.x, S); }
// end of day d (capacity scan) fused with the start of day d+1 (stats row, queue, tracing, ...): one launch less per day
__global__ void __launch_bounds__(PRE_THREADS) k_between(Eng G) {
    __shared__ SmemSmall S;
    post_body(G, blockIdx.x, S);
    __syncthreads();
    pre_body(G, blockIdx.x, S);
}

// ---------------------------------------------------------------- initial population condition
// Population.set_initial_state, main.pyx:1452-1516 (Context.__init__ :1780-1781: day 0, testing still NO_TESTING).
// One-time, order-dependent setup of a few thousand people drawn WITH replacement: lane 0 of one CTA per replica replays
// the reference's loop literally (see apply_initial_state in the oracle for the quirks that are kept).
struct Ipc { int32_t dead, in_icu, in_ward, confirmed, incubating, ill, recovered; };

__device__ void init_remove(const Eng &G, RepCtr *c, size_t base, int32_t a, int age, bool dies) {   // person_recover / person_die
    uint32_t h = G.hot[base + a];
    count_add(c, RB_A_INFECTED, age, -1);
    if (dies) { count_add(c, RB_A_DEAD, age, 1); if (H_SEV(h) == RB_FATAL) count_add(c, RB_A_NON_HOSPITAL_DEATHS, age, 1); }
    else count_add(c, RB_A_RECOVERED, age, 1);
    G.hot[base + a] = H_SET_STATE(h, dies ? RB_DEAD : RB_RECOVERED) & ~(H_LIST | H_FRESH);
}
__device__ bool init_dies_without_care(const Eng &G, RepCtr *c, int32_t a, uint32_t h) {   // dies_in_hospital(care_available=False)
    const uint32_t sev = H_SEV(h);
    if (sev == RB_FATAL) return true;
    const rb_variant *v = &G.variants[H_VAR(h)];
    const float ch = sev == RB_CRITICAL ? v->p_icu_death_no_beds : (sev == RB_SEVERE ? v->p_hospital_death_no_beds : 0.0f);
    u32x4 x = philox(c->seed, (uint32_t)a, (uint32_t)c->day, PU_NOBED, 0);
    return chance(u01d(x.x, x.y), ch);
}
__device__ void init_hospitalize(const Eng &G, RepCtr *c, size_t base, int32_t a, int age) {   // person_hospitalize, main.pyx:321-338
    uint32_t h = G.hot[base + a];
    if (!(h & H_DET)) { h |= H_DET; count_add(c, RB_A_DETECTED, age, 1); count_add(c, RB_A_ALL_DETECTED, age, 1); G.hot[base + a] = h; }
    if (c->avail_beds == 0) { init_remove(G, c, base, a, age, init_dies_without_care(G, c, a, h)); return; }
    c->avail_beds -= 1;
    G.hot[base + a] = H_SET_DL(H_SET_STATE(h, RB_HOSPITALIZED), (G.rec[base + a].cold >> 16) & 255u) & ~H_FRESH;
    count_add(c, RB_A_IN_WARD, age, 1);
}
__device__ void init_to_icu(const Eng &G, RepCtr *c, size_t base, int32_t a, int age) {   // person_transfer_to_icu, main.pyx:341-351
    uint32_t h = G.hot[base + a];
    c->avail_beds += 1;
    if (c->avail_icu == 0) {
        if (init_dies_without_care(G, c, a, h)) { count_add(c, RB_A_IN_WARD, age, -1); init_remove(G, c, base, a, age, true); return; }
    } else c->avail_icu -= 1;
    G.hot[base + a] = H_SET_DL(H_SET_STATE(h, RB_IN_ICU), (G.rec[base + a].cold >> 24) & 255u) & ~H_FRESH;
    count_add(c, RB_A_IN_WARD, age, -1); count_add(c, RB_A_IN_ICU, age, 1); count_add(c, RB_A_CUM_ICU, age, 1);
}
__device__ void init_become_ill(const Eng &G, RepCtr *c, size_t base, int32_t a) {   // person_become_ill, main.pyx:284-291 (nobody seeks testing yet)
    uint32_t h = G.hot[base + a];
    const uint32_t sev = H_SEV(h);
    const rb_variant *v = &G.variants[H_VAR(h)];
    const int day = c->day;
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
    G.rec[base + a].cold = (G.rec[base + a].cold & 0xffffu) | (wd << 16) | (ud << 24);     // assigned, not OR-ed: the person may have been drawn before
    G.hot[base + a] = H_SET_DL(H_SET_STATE(h, RB_ILLNESS), dl) & ~H_FRESH;
}

__global__ void k_initial_state(Eng G, Ipc P) {
    if (threadIdx.x != 0) return;
    const int r = blockIdx.x;
    RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const int32_t were_ill = P.dead + P.recovered + P.in_icu + P.in_ward + P.ill, were_incubating = were_ill + P.incubating;
    const int32_t i_incubating = P.incubating, i_rws = i_incubating + (were_incubating - were_ill);
    const int32_t i_ill_at_home = i_rws + P.ill, i_dead = i_ill_at_home + P.dead, i_in_icu = i_dead + P.in_icu, i_in_ward = i_in_icu + P.in_ward;
    for (int32_t i = 0; i < were_incubating; i++) {
        u32x4 x = philox(c->seed, (uint32_t)i, 0u, PU_INIT, 0);
        const int32_t a = (int32_t)(x.x % (uint32_t)G.N);           // get_random_person, main.pyx:1518-1523
        const int age = age_of(G, a);
        // a person drawn before is infected again, exactly as the reference does; person_infect (main.pyx:209-235) resets
        // state, severity and the day counter but leaves was_detected alone
        const uint32_t was_detected = G.hot[base + a] & H_DET;
        device_infect(G, r, c, a, -1, 0u, 0, 0, true);
        if (was_detected) G.hot[base + a] |= H_DET;
        if (i < i_incubating) continue;                              // still incubating: waits one day like any same-day infection
        if (i < i_rws) { init_remove(G, c, base, a, age, false); continue; }
        init_become_ill(G, c, base, a);
        if (i < i_ill_at_home) continue;
        if (i < i_dead) { init_remove(G, c, base, a, age, true); continue; }
        if (i < i_in_icu) { init_hospitalize(G, c, base, a, age); init_to_icu(G, c, base, a, age); continue; }
        if (i < i_in_ward) { init_hospitalize(G, c, base, a, age); continue; }
        init_remove(G, c, base, a, age, false);
    }
    for (int age = 0; age < 100 && age < G.n_ages; age++) c->counts[RB_A_ALL_DETECTED][age] = 0;
    for (int32_t i = 0; i < P.confirmed; i++) c->counts[RB_A_ALL_DETECTED][(100 + i) % 100] += 1;
}

// ---------------------------------------------------------------- misc kernels
__global__ void k_init(Eng G) {
    const int r = blockIdx.y;
    const size_t base = (size_t)r * G.Npad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < G.Npad; i += gridDim.x * blockDim.x) {
        // padding words beyond N are marked RECOVERED+included so that the sweep skips them
        G.hot[base + i] = i < G.N ? 0u : (RB_RECOVERED | H_INCL);
        AgentRec z; z.winner = KEY_IDLE; z.infector = -1; z.first_child = -1; z.next_sib = -1; z.inf_key = 0; z.cold = 0; z.vacc_day = -1; z.pad = 0;
        G.rec[base + i] = z;
    }
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < G.sus_words; w += gridDim.x * blockDim.x) {
        int first = w * 32;
        uint32_t m = first + 32 <= G.N ? 0xffffffffu : (first >= G.N ? 0u : ((1u << (G.N - first)) - 1u));
        G.sus[(size_t)r * G.sus_words + w] = m;
        G.act[(size_t)r * G.sus_words + w] = 0u;
    }
}

__global__ void k_snapshot(Eng G) {
    __shared__ int32_t srow[RB_N_ATTRS * 16 + RB_N_SCALARS];
    write_stats_row(G, blockIdx.x, &G.ctr[blockIdx.x], srow);
}

// Per-day metric aggregation across the ensemble: sum and sum of squares over replicas of every stats column.
__global__ void k_moments(Eng G, int day0, double *out_sum, double *out_sq) {
    const int d = blockIdx.x;
    for (int col = threadIdx.x; col < G.row_len; col += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < G.R; r++) {
            const double v = (double)G.stats[((size_t)r * (G.max_days + 1) + day0 + d) * G.row_len + col];
            s1 += v; s2 += v * v;
        }
        out_sum[(size_t)d * G.row_len + col] = s1; out_sq[(size_t)d * G.row_len + col] = s2;
    }
}

// Context.sample, main.pyx:2047-2101
__global__ void k_sample(Eng G, int what, int age, int severity, int n, int epoch, int32_t *out) {
    const rb_variant *v = &G.variants[0];
    uint32_t seed = G.ctr[0].seed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t pu = PU_SAMPLE | ((uint32_t)what << 8);
        int res;
        if (what == 0) {
            u32x4 x = philox(seed, (uint32_t)i, (uint32_t)age, pu, 0);
            double u = u01d(x.x, x.y);
            const double *cdf = G.tables[epoch]->ncdf[age][0];
            int lo = 0, hi = 100;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (u < cdf[mid]) hi = mid; else lo = mid + 1; }
            res = lo;
        } else if (what == 1) {
            u32x4 x = philox(seed, (uint32_t)i, (uint32_t)age, pu, 0);
            res = symptom_severity(v, age, u01f(x.x), false);
        } else if (what == 2) {
            res = round_to_int(gamma_f(seed, (uint32_t)i, (uint32_t)age, pu, v->incubation_kappa, v->incubation_theta));
        } else {
            float T = severity == RB_FATAL ? gamma_f(seed, (uint32_t)i, (uint32_t)age, pu, v->onset_death_kappa, v->onset_death_theta)
                                           : gamma_f(seed, (uint32_t)i, (uint32_t)age, pu, v->onset_recovery_kappa, v->onset_recovery_theta);
            float f = 0.0f;
            if (what == 3) { f = T; if (severity != RB_ASYMPTOMATIC && severity != RB_MILD) f = f * v->ratio_before_hospitalisation; }
            else if (what == 4) { if (severity == RB_SEVERE) f = T * (1.0f - v->ratio_before_hospitalisation); else if (severity >= RB_CRITICAL) f = T * v->ratio_in_ward; }
            else if (what == 5) { if (severity >= RB_CRITICAL) f = ((1.0f - v->ratio_in_ward) - v->ratio_before_hospitalisation) * T; }
            else f = T;
            res = round_to_int(f);
        }
        out[i] = res;
    }
}

// ================================================================ host side / C-ABI
static thread_local char g_err[512];
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return 1; } } while (0)

struct rb_engine {
    rb_config cfg;
    Eng G;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    std::vector<void *> allocs;
    std::vector<DevTable *> tables;     // host copy of the device pointer table
    DevTable **d_tables;
    int n_table_slots;
    rb_day_params *d_sched;
    std::vector<rb_day_params> h_sched;
    std::vector<int32_t> age_start, age_counts;
    std::vector<rb_variant> h_variants;
    int32_t day;
    float last_ms;
    int64_t launches;
    int sweep_blocks, list_blocks, resolve_blocks;
    cudaGraphExec_t graph[2];
    bool have_graphs;
    ncclComm_t comm;                    // population-sharded mode
    int merge_blocks;
    Ipc ipc; bool has_ipc;              // initial population condition, re-applied by rb_reset
};

// ---------------------------------------------------------------- NCCL, bound at run time
struct NcclApi {
    void *dl;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.dl) return 0;
    // reuse a libnccl the process already mapped (torch ships one with the same soname), else load the system one
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { snprintf(g_err, sizeof g_err, "cannot load libnccl.so.2: %s", dlerror()); return 1; }
    g_nccl.GetUniqueId = (ncclResult_t(*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (ncclResult_t(*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (ncclResult_t(*)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(ncclResult_t))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy || !g_nccl.GetErrorString) {
        snprintf(g_err, sizeof g_err, "libnccl.so.2 lacks a required symbol"); return 1;
    }
    g_nccl.dl = h;
    return 0;
}
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { snprintf(g_err, sizeof g_err, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); return 1; } } while (0)


template <typename T> static int dalloc(rb_engine *e, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t err = cudaMalloc(&q, n * sizeof(T));
    if (err != cudaSuccess) { snprintf(g_err, sizeof g_err, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(err)); return 1; }
    e->allocs.push_back(q);
    *p = (T *)q;
    return 0;
}
static uint32_t pow2_at_least(uint64_t x) { uint32_t p = 1024; while (p < x) p <<= 1; return p; }

extern "C" const char *rb_last_error(void) { return g_err; }

static int init_counters(rb_engine *e, uint32_t seed) {
    const rb_config *cfg = &e->cfg;
    const int R = cfg->n_replicas;
    std::vector<RepCtr> hc(R);
    memset(hc.data(), 0, sizeof(RepCtr) * R);
    for (int r = 0; r < R; r++) {
        RepCtr &c = hc[r];
        for (int a = 0; a < cfg->n_ages; a++) c.counts[RB_A_SUSCEPTIBLE][a] = e->age_counts[a];
        c.beds = c.avail_beds = cfg->hospital_beds; c.icu = c.avail_icu = cfg->icu_units;
        c.p_successful_tracing = 1.0f;
        c.seed = seed + (uint32_t)r;
        u32x4 k = philox(c.seed, 0, 0, PU_PERM, 0);
        c.fkey[0] = k.x; c.fkey[1] = k.y; c.fkey[2] = k.z; c.fkey[3] = k.w;
        for (int i = 0; i < RB_MAX_VACC; i++) c.vacc_cursor[i] = -2;
    }
    CK(cudaMemcpyAsync(e->G.ctr, hc.data(), sizeof(RepCtr) * R, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" void rb_destroy(rb_engine *e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaStreamSynchronize(e->stream);
    if (e->have_graphs) { cudaGraphExecDestroy(e->graph[0]); cudaGraphExecDestroy(e->graph[1]); }
    if (e->comm) g_nccl.CommDestroy(e->comm);
    for (void *p : e->allocs) cudaFree(p);
    cudaEventDestroy(e->ev0); cudaEventDestroy(e->ev1);
    cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" int rb_create(const rb_config *cfg, const int32_t *age_counts, const int32_t *group_of_age,
                         const rb_variant *variants, const int32_t *import_lo, const int32_t *import_hi,
                         const float *import_cum, rb_engine **out) {
    if (cfg->n_ages > RB_MAX_AGES || cfg->n_variants > RB_MAX_VARIANTS || cfg->n_import_classes > RB_MAX_IMPORT_CLASSES ||
        cfg->n_groups > 16 || cfg->n_replicas < 1 || cfg->n_agents < 1) {
        snprintf(g_err, sizeof g_err, "config exceeds compiled limits"); return 1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        snprintf(g_err, sizeof g_err, "no CUDA device: reina_b200 has no CPU fallback"); return 2;
    }
    CK(cudaSetDevice(cfg->device));
    rb_engine *e = new rb_engine();
    e->cfg = *cfg; e->day = 0; e->last_ms = 0; e->launches = 0; e->have_graphs = false; e->comm = nullptr; e->merge_blocks = 1;
    e->has_ipc = false;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&e->ev0)); CK(cudaEventCreate(&e->ev1));
    Eng &G = e->G;
    memset(&G, 0, sizeof G);
    const int N = cfg->n_agents, R = cfg->n_replicas;
    G.N = N; G.Npad = (N + 3) & ~3; G.n_ages = cfg->n_ages; G.n_groups = cfg->n_groups; G.n_variants = cfg->n_variants;
    G.R = R; G.max_days = cfg->max_days; G.row_len = RB_N_ATTRS * cfg->n_groups + RB_N_SCALARS;
    G.n_import_classes = cfg->n_import_classes;
    G.rank = 0; G.nranks = 1;
    int bits = 1; while ((1u << bits) < (uint32_t)N) bits++;
    G.fhalf = (bits + 1) / 2;
    e->age_start.resize(cfg->n_ages + 1);
    int64_t tot = 0;
    for (int a = 0; a < cfg->n_ages; a++) { e->age_start[a] = (int32_t)tot; tot += age_counts[a]; }
    e->age_start[cfg->n_ages] = (int32_t)tot;
    if (tot != N) { snprintf(g_err, sizeof g_err, "age_counts sum %lld != n_agents %d", (long long)tot, N); delete e; return 1; }
    float cc = cfg->contact_capacity > 0 ? cfg->contact_capacity : 1.0f;
    G.cap_items = pow2_at_least((uint64_t)((double)N * cc) + 4096);
    G.cap_succ = pow2_at_least((uint64_t)N / 8 + 4096);
    G.cap_events = pow2_at_least((uint64_t)N / 16 + 2048);
    G.cap_queue = pow2_at_least((uint64_t)N / 8 + 2048);
    const size_t RN = (size_t)R * G.Npad;
    G.sus_words = ((G.Npad + 31) / 32 + 32 + 3) & ~3;     // multiple of 4 words: the sweep reads the bitmaps 16 bytes at a time
    if (dalloc(e, &G.hot, RN) || dalloc(e, &G.rec, RN) ||
        dalloc(e, &G.sus, (size_t)R * G.sus_words) || dalloc(e, &G.act, (size_t)R * G.sus_words) || dalloc(e, &G.items, (size_t)R * G.cap_items) || dalloc(e, &G.succ, (size_t)R * G.cap_succ) ||
        dalloc(e, &G.ev_key, (size_t)R * G.cap_events) || dalloc(e, &G.ev_agent, (size_t)R * G.cap_events) ||
        dalloc(e, &G.q_key, (size_t)R * 2 * G.cap_queue) || dalloc(e, &G.q_agent, (size_t)R * 2 * G.cap_queue) ||
        dalloc(e, &G.ctr, (size_t)R) || dalloc(e, &G.stats, (size_t)R * (cfg->max_days + 1) * G.row_len) ||
        dalloc(e, &e->d_sched, (size_t)cfg->max_days + 1)) { rb_destroy(e); return 1; }
    G.sched = e->d_sched;
    e->h_sched.assign(cfg->max_days + 1, rb_day_params());
    e->n_table_slots = 1024;
    if (dalloc(e, &e->d_tables, (size_t)e->n_table_slots)) { rb_destroy(e); return 1; }
    CK(cudaMemset(e->d_tables, 0, sizeof(DevTable *) * e->n_table_slots));
    G.tables = e->d_tables;
    e->tables.assign(e->n_table_slots, nullptr);
    rb_variant *dv; int32_t *d_as, *d_ga, *d_ilo, *d_ihi; float *d_icum; uint8_t *d_ab;
    if (dalloc(e, &dv, (size_t)cfg->n_variants) || dalloc(e, &d_as, (size_t)cfg->n_ages + 1) || dalloc(e, &d_ga, (size_t)cfg->n_ages) ||
        dalloc(e, &d_ilo, (size_t)RB_MAX_IMPORT_CLASSES) || dalloc(e, &d_ihi, (size_t)RB_MAX_IMPORT_CLASSES) ||
        dalloc(e, &d_icum, (size_t)RB_MAX_IMPORT_CLASSES) || dalloc(e, &d_ab, (size_t)(G.Npad >> 10) + 2)) { rb_destroy(e); return 1; }
    {
        std::vector<uint8_t> ab((size_t)(G.Npad >> 10) + 2);
        int age = 0;
        for (size_t b = 0; b < ab.size(); b++) {
            int64_t a = (int64_t)b << 10; if (a > N - 1) a = N - 1;
            while (age < cfg->n_ages - 1 && e->age_start[age + 1] <= a) age++;
            ab[b] = (uint8_t)age;
        }
        CK(cudaMemcpy(d_ab, ab.data(), ab.size(), cudaMemcpyHostToDevice));
    }
    e->h_variants.assign(variants, variants + cfg->n_variants);
    CK(cudaMemcpy(dv, variants, sizeof(rb_variant) * cfg->n_variants, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_as, e->age_start.data(), sizeof(int32_t) * (cfg->n_ages + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ga, group_of_age, sizeof(int32_t) * cfg->n_ages, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ilo, import_lo, sizeof(int32_t) * cfg->n_import_classes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ihi, import_hi, sizeof(int32_t) * cfg->n_import_classes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_icum, import_cum, sizeof(float) * cfg->n_import_classes, cudaMemcpyHostToDevice));
    G.age_blk = d_ab; G.variants = dv; G.age_start = d_as; G.group_of_age = d_ga; G.import_lo = d_ilo; G.import_hi = d_ihi; G.import_cum = d_icum;
    e->age_counts.assign(age_counts, age_counts + cfg->n_ages);
    if (init_counters(e, cfg->seed)) { rb_destroy(e); return 1; }
    CK(cudaMemset(G.stats, 0, sizeof(int32_t) * (size_t)R * (cfg->max_days + 1) * G.row_len));
    CK(cudaMemset(e->d_sched, 0, sizeof(rb_day_params) * ((size_t)cfg->max_days + 1)));
    // launch geometry: grid-stride kernels sized in multiples of the SM count
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, cfg->device));
    int sms = prop.multiProcessorCount;
    // the sweep fills the GPU exactly once: SW_CTAS_PER_SM resident CTAs per SM, shared out over the replicas (a grid a
    // little larger than one wave would run its tail on a nearly empty GPU)
    int want = (G.sus_words / 4 + SW_WARPS * 32 - 1) / (SW_WARPS * 32);
    int per_rep = sms * SW_CTAS_PER_SM / R; if (per_rep < 1) per_rep = 1;
    e->sweep_blocks = want < per_rep ? want : per_rep; if (e->sweep_blocks < 1) e->sweep_blocks = 1;
    CK(cudaFuncSetAttribute(k_sweep, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));   // 8 x 24 KB per SM
    e->list_blocks = sms * EX_CTAS_PER_SM / R; if (e->list_blocks < 2) e->list_blocks = 2;
    // k_resolve is a chain of dependent scattered accesses per infection: enough threads for one pass over the day's list
    e->resolve_blocks = (int)((G.N / 128 + 255) / 256); if (e->resolve_blocks < e->list_blocks) e->resolve_blocks = e->list_blocks;
    { int cap = sms * 16 / R; if (cap < 64) cap = 64; if (e->resolve_blocks > cap) e->resolve_blocks = cap; }
    k_init<<<dim3(e->sweep_blocks, R), 256, 0, e->stream>>>(G); e->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    *out = e;
    return 0;
}

extern "C" int rb_reset(rb_engine *e, uint32_t seed) {
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    e->cfg.seed = seed;
    e->day = 0;
    if (init_counters(e, seed)) return 1;
    k_init<<<dim3(e->sweep_blocks, e->G.R), 256, 0, e->stream>>>(e->G); e->launches++;
    if (e->has_ipc) { k_initial_state<<<e->G.R, 32, 0, e->stream>>>(e->G, e->ipc); e->launches++; }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int rb_set_initial_state(rb_engine *e, const int32_t *ipc7) {
    CK(cudaSetDevice(e->cfg.device));
    if (e->day != 0 || e->has_ipc) { snprintf(g_err, sizeof g_err, "rb_set_initial_state: once, before the first step"); return 1; }
    for (int i = 0; i < 7; i++) if (ipc7[i] < 0) { snprintf(g_err, sizeof g_err, "rb_set_initial_state: negative count"); return 1; }
    e->ipc.dead = ipc7[0]; e->ipc.in_icu = ipc7[1]; e->ipc.in_ward = ipc7[2]; e->ipc.confirmed = ipc7[3];
    e->ipc.incubating = ipc7[4]; e->ipc.ill = ipc7[5]; e->ipc.recovered = ipc7[6];
    e->has_ipc = true;
    k_initial_state<<<e->G.R, 32, 0, e->stream>>>(e->G, e->ipc); e->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int rb_set_contact_table(rb_engine *e, int32_t epoch, const int32_t *n_rows, const double *cum_p,
                                    const int32_t *age_lo, const int32_t *age_hi, const uint8_t *place,
                                    const float *mask_p, const double *nr_contacts, const double *ncontact_cdf) {
    if (epoch < 0 || epoch >= e->n_table_slots) { snprintf(g_err, sizeof g_err, "table epoch out of range"); return 1; }
    CK(cudaSetDevice(e->cfg.device));
    DevTable *h = new DevTable();
    memset(h, 0, sizeof *h);
    for (int age = 0; age < e->cfg.n_ages; age++) {
        h->n_rows[age] = n_rows[age];
        h->nr_contacts[age] = (float)nr_contacts[age];
        memcpy(h->ncdf[age], ncontact_cdf + (size_t)age * 2 * RB_NCDF, sizeof(double) * 2 * RB_NCDF);
        for (int i = 0; i < n_rows[age]; i++) {
            int k = age * RB_MAX_ROWS + i;
            h->cum_p[age][i] = cum_p[k];
            {   // exact: scaling by 2^24 is exact in double, and for integer k, k < x  <=>  k < ceil(x)
                double x = cum_p[k] * 16777216.0, cx = (double)(uint64_t)x;
                if (cx < x) cx += 1.0;
                h->cum24[age][i] = cx >= 4294967295.0 ? 0xffffffffu : (uint32_t)cx;
            }
            h->start[age][i] = e->age_start[age_lo[k]];
            h->size[age][i] = e->age_start[age_hi[k] + 1] - e->age_start[age_lo[k]];
            h->place[age][i] = place[k];
            h->lo_age[age][i] = (uint8_t)age_lo[k]; h->hi_age[age][i] = (uint8_t)age_hi[k];
            bool uni = true;
            for (int v = 0; v < e->cfg.n_variants; v++)
                for (int g = age_lo[k]; g <= age_hi[k]; g++)
                    if (e->h_variants[v].tab[RB_T_SUSCEPTIBILITY][g] != e->h_variants[v].tab[RB_T_SUSCEPTIBILITY][age_lo[k]]) uni = false;
            h->susc_uniform[age][i] = uni ? 1 : 0;
            h->mask_p[age][i] = mask_p[k];
        }
    }
    for (int age = 0; age < e->cfg.n_ages; age++)
        for (int cls = 0; cls < 2; cls++)
            for (int b = 0, k = 0; b < 256; b++) {
                const int limit = cls ? 5 : 100;
                while (k < limit && !(h->ncdf[age][cls][k] > (double)b / 256.0)) k++;
                h->nguide[age][cls][b] = (uint8_t)k;
            }
    for (int age = 0; age < e->cfg.n_ages; age++)
        for (int b = 0, i = 0; b < 1024; b++) {          // cum_p is non-decreasing: the start row only moves forward
            while (i < n_rows[age] - 1 && !(h->cum_p[age][i] > (double)b / 1024.0)) i++;
            h->guide[age][b] = (uint8_t)i;
        }
    DevTable *d = e->tables[epoch];
    if (!d) { if (dalloc(e, &d, 1)) { delete h; return 1; } e->tables[epoch] = d; }
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(d, h, sizeof(DevTable), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_tables + epoch, &d, sizeof(DevTable *), cudaMemcpyHostToDevice));
    delete h;
    return 0;
}

extern "C" int rb_set_schedule(rb_engine *e, int32_t day0, int32_t n, const rb_day_params *params) {
    if (day0 < 0 || day0 + n > e->cfg.max_days + 1) { snprintf(g_err, sizeof g_err, "schedule out of range"); return 1; }
    CK(cudaSetDevice(e->cfg.device));
    memcpy(e->h_sched.data() + day0, params, sizeof(rb_day_params) * n);
    CK(cudaMemcpyAsync(e->d_sched + day0, e->h_sched.data() + day0, sizeof(rb_day_params) * n, cudaMemcpyHostToDevice, e->stream));
    return 0;
}

// One "segment" = the grid kernels of day d followed by the fused day boundary d -> d+1.
static void launch_segment(rb_engine *e, cudaStream_t st) {
    const Eng &G = e->G;
    k_sweep<<<dim3(e->sweep_blocks, G.R), SW_THREADS, 0, st>>>(G);
    k_expose<<<dim3(e->list_blocks, G.R), EX_THREADS, 0, st>>>(G);
    k_resolve<true><<<dim3(e->resolve_blocks, G.R), 256, 0, st>>>(G);
    k_between<<<G.R, PRE_THREADS, 0, st>>>(G);
}

// No kernel takes the day as an argument (each replica carries its own day counter and reads the schedule from
// device memory), so a captured graph of GRAPH_DAYS segments is replayed for any stretch of days.
#define GRAPH_DAYS 16
static int build_graphs(rb_engine *e) {
    for (int which = 0; which < 2; which++) {
        int nseg = which == 0 ? GRAPH_DAYS : 1;
        cudaGraph_t g;
        CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < nseg; i++) launch_segment(e, e->stream);
        CK(cudaStreamEndCapture(e->stream, &g));
        CK(cudaGraphInstantiate(&e->graph[which], g, 0));
        CK(cudaGraphDestroy(g));
    }
    e->have_graphs = true;
    return 0;
}

// ---------------------------------------------------------------- population-sharded mode
extern "C" int rb_shard_unique_id(uint8_t *out128) {
    if (load_nccl()) return 1;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out128, &id, 128);
    return 0;
}

extern "C" int rb_shard_init(rb_engine *e, int32_t rank, int32_t nranks, const uint8_t *uid128, float exchange_capacity) {
    if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) { snprintf(g_err, sizeof g_err, "bad rank %d / %d", rank, nranks); return 1; }
    if (e->G.R != 1) { snprintf(g_err, sizeof g_err, "population-sharded mode runs one replica (n_replicas = %d)", e->G.R); return 1; }
    if (e->day != 0 || e->comm) { snprintf(g_err, sizeof g_err, "rb_shard_init must be the first call after rb_create"); return 1; }
    if (load_nccl()) return 1;
    CK(cudaSetDevice(e->cfg.device));
    ncclUniqueId id; memcpy(&id, uid128, 128);
    NK(g_nccl.CommInitRank(&e->comm, nranks, id, rank));
    Eng &G = e->G;
    // message capacities: this rank's share of the agents; a day's state changes / transmissions / tests / capacity
    // events are small fractions of it (peak day of the reference epidemic: 0.9 % / 0.5 % / 0.17 % / 0.08 % of the
    // agents; ~2.3x headroom each, exchange_capacity scales them, overflow is a loud RB_OTHER_FAILURE)
    const double share = (double)G.N / nranks * (exchange_capacity > 0 ? exchange_capacity : 1.0);
    G.xcap_upd = (uint32_t)(share / 48) + 4096;
    G.xcap_succ = (uint32_t)(share / 96) + 4096;
    G.xcap_q = (uint32_t)(share / 256) + 2048;
    G.xcap_ev = (uint32_t)(share / 512) + 2048;
    G.xslot = xslot_bytes(G.xcap_q, G.xcap_ev, G.xcap_upd, G.xcap_succ);
    if (dalloc(e, &G.xbuf, G.xslot * nranks)) return 1;
    CK(cudaMemset(G.xbuf, 0, G.xslot * nranks));
    G.rank = rank; G.nranks = nranks;
    int sb = (e->sweep_blocks + nranks - 1) / nranks; if (sb < 1) sb = 1;
    e->sweep_blocks = sb;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, e->cfg.device));
    e->merge_blocks = prop.multiProcessorCount * 2;
    return 0;
}

extern "C" int32_t rb_shard_rank(rb_engine *e) { return e->G.rank; }
extern "C" int32_t rb_shard_nranks(rb_engine *e) { return e->G.nranks; }
extern "C" int64_t rb_shard_message_bytes(rb_engine *e) { return e->comm ? (int64_t)e->G.xslot : 0; }

// One simulated day in sharded mode: sweep and contacts over the owned stripes, ONE all-gather of the ranks' messages,
// then merge / resolve / day boundary replicated on every rank.
static int launch_day_sharded(rb_engine *e, bool last) {
    const Eng &G = e->G;
    cudaStream_t st = e->stream;
    k_sweep<<<dim3(e->sweep_blocks, 1), SW_THREADS, 0, st>>>(G);
    k_expose<<<dim3(e->list_blocks, 1), EX_THREADS, 0, st>>>(G);
    NK(g_nccl.AllGather(G.xbuf + (size_t)G.rank * G.xslot, G.xbuf, G.xslot, ncclChar, e->comm, st));
    k_merge<<<e->merge_blocks, 256, 0, st>>>(G);
    if (last) { k_resolve<false><<<dim3(e->resolve_blocks, 1), 256, 0, st>>>(G); k_post<<<1, PRE_THREADS, 0, st>>>(G); }
    else { k_resolve<true><<<dim3(e->resolve_blocks, 1), 256, 0, st>>>(G); k_between<<<1, PRE_THREADS, 0, st>>>(G); }
    e->launches += 5;
    return 0;
}

extern "C" int rb_step(rb_engine *e, int32_t n_days) {
    CK(cudaSetDevice(e->cfg.device));
    if (n_days <= 0) return 0;
    if (e->day + n_days > e->cfg.max_days) { snprintf(g_err, sizeof g_err, "max_days exceeded"); return 1; }
    for (int d = 0; d < n_days; d++) {
        int ep = e->h_sched[e->day + d].table_epoch;
        if (ep < 0 || ep >= e->n_table_slots || !e->tables[ep]) { snprintf(g_err, sizeof g_err, "contact table %d not set", ep); return 1; }
    }
    if (!e->comm && !e->have_graphs && build_graphs(e)) return 1;
    const Eng &G = e->G;
    const int R = G.R;
    if (e->comm) {
        CK(cudaEventRecord(e->ev0, e->stream));
        k_pre<<<1, PRE_THREADS, 0, e->stream>>>(G); e->launches++;
        for (int d = 0; d < n_days; d++) if (launch_day_sharded(e, d == n_days - 1)) return 1;
        CK(cudaEventRecord(e->ev1, e->stream));
        CK(cudaGetLastError());
        e->day += n_days;
        return 0;
    }
    CK(cudaEventRecord(e->ev0, e->stream));
    k_pre<<<R, PRE_THREADS, 0, e->stream>>>(G); e->launches++;
    int mid = n_days - 1;
    while (mid >= GRAPH_DAYS) { CK(cudaGraphLaunch(e->graph[0], e->stream)); mid -= GRAPH_DAYS; e->launches += 4 * GRAPH_DAYS; }
    while (mid > 0) { CK(cudaGraphLaunch(e->graph[1], e->stream)); mid -= 1; e->launches += 4; }
    k_sweep<<<dim3(e->sweep_blocks, R), SW_THREADS, 0, e->stream>>>(G);
    k_expose<<<dim3(e->list_blocks, R), EX_THREADS, 0, e->stream>>>(G);
    k_resolve<false><<<dim3(e->resolve_blocks, R), 256, 0, e->stream>>>(G);
    k_post<<<R, PRE_THREADS, 0, e->stream>>>(G);
    e->launches += 4;
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaGetLastError());
    e->day += n_days;
    return 0;
}

extern "C" int rb_step_profiled(rb_engine *e, int32_t n_days, float *ms_per_kernel) {
    CK(cudaSetDevice(e->cfg.device));
    if (e->day + n_days > e->cfg.max_days) { snprintf(g_err, sizeof g_err, "max_days exceeded"); return 1; }
    if (e->comm) { snprintf(g_err, sizeof g_err, "rb_step_profiled: not available in population-sharded mode"); return 1; }
    const Eng &G = e->G;
    const int R = G.R;
    std::vector<cudaEvent_t> ev((size_t)n_days * 6);
    for (auto &x : ev) CK(cudaEventCreate(&x));
    for (int d = 0; d < n_days; d++) {
        cudaEvent_t *v = &ev[(size_t)d * 6];
        CK(cudaEventRecord(v[0], e->stream));
        k_pre<<<R, PRE_THREADS, 0, e->stream>>>(G); CK(cudaEventRecord(v[1], e->stream));
        k_sweep<<<dim3(e->sweep_blocks, R), SW_THREADS, 0, e->stream>>>(G); CK(cudaEventRecord(v[2], e->stream));
        k_expose<<<dim3(e->list_blocks, R), EX_THREADS, 0, e->stream>>>(G); CK(cudaEventRecord(v[3], e->stream));
        k_resolve<false><<<dim3(e->resolve_blocks, R), 256, 0, e->stream>>>(G); CK(cudaEventRecord(v[4], e->stream));
        k_post<<<R, PRE_THREADS, 0, e->stream>>>(G); CK(cudaEventRecord(v[5], e->stream));
        e->launches += 5;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < RB_N_KERNELS; k++) ms_per_kernel[k] = 0;
    for (int d = 0; d < n_days; d++)
        for (int k = 0; k < RB_N_KERNELS; k++) { float ms = 0; CK(cudaEventElapsedTime(&ms, ev[(size_t)d * 6 + k], ev[(size_t)d * 6 + k + 1])); ms_per_kernel[k] += ms; }
    for (auto &x : ev) cudaEventDestroy(x);
    e->day += n_days;
    return 0;
}

extern "C" int rb_sync(rb_engine *e) {
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->last_ms = ms; else cudaGetLastError();
    return 0;
}

extern "C" int32_t rb_day(rb_engine *e) { return e->day; }
extern "C" void rb_debug_flag(rb_engine *e, int32_t v) { e->G.dbg = v; }
extern "C" int rb_debug_phase_cycles(rb_engine *e, int32_t replica, long long *out16) {
    cudaSetDevice(e->cfg.device); cudaStreamSynchronize(e->stream);
    RepCtr c; if (cudaMemcpy(&c, &e->G.ctr[replica], sizeof c, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    memcpy(out16, c.dbg_t, sizeof c.dbg_t); return 0;
}
extern "C" int32_t rb_row_len(rb_engine *e) { return e->G.row_len; }
extern "C" float rb_last_step_ms(rb_engine *e) { return e->last_ms; }
extern "C" int64_t rb_launch_count(rb_engine *e) { return e->launches; }

extern "C" int rb_snapshot(rb_engine *e) {
    CK(cudaSetDevice(e->cfg.device));
    k_snapshot<<<e->G.R, 256, 0, e->stream>>>(e->G); e->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int rb_read_stats(rb_engine *e, int32_t day0, int32_t n, int32_t *out) {
    CK(cudaSetDevice(e->cfg.device));
    if (day0 < 0 || day0 + n > e->cfg.max_days + 1) { snprintf(g_err, sizeof g_err, "stats range"); return 1; }
    const Eng &G = e->G;
    CK(cudaMemcpy2DAsync(out, sizeof(int32_t) * (size_t)n * G.row_len,
                         G.stats + (size_t)day0 * G.row_len, sizeof(int32_t) * (size_t)(G.max_days + 1) * G.row_len,
                         sizeof(int32_t) * (size_t)n * G.row_len, G.R, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int rb_read_moments(rb_engine *e, int32_t day0, int32_t n, double *sum, double *sumsq) {
    CK(cudaSetDevice(e->cfg.device));
    if (day0 < 0 || n < 1 || day0 + n > e->cfg.max_days + 1) { snprintf(g_err, sizeof g_err, "stats range"); return 1; }
    const size_t cnt = (size_t)n * e->G.row_len;
    double *d; CK(cudaMalloc(&d, sizeof(double) * 2 * cnt));
    k_moments<<<n, 160, 0, e->stream>>>(e->G, day0, d, d + cnt); e->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(sum, d, sizeof(double) * cnt, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(sumsq, d + cnt, sizeof(double) * cnt, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d);
    return 0;
}

extern "C" int rb_read_per_age(rb_engine *e, int32_t replica, int32_t attr, int32_t *out) {
    CK(cudaSetDevice(e->cfg.device));
    if (replica < 0 || replica >= e->G.R || attr < 0 || attr >= RB_N_ATTRS) { snprintf(g_err, sizeof g_err, "bad replica/attr"); return 1; }
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out, &e->G.ctr[replica].counts[attr][0], sizeof(int32_t) * e->cfg.n_ages, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int rb_problem(rb_engine *e, int32_t *out) {
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    for (int r = 0; r < e->G.R; r++) CK(cudaMemcpy(out + r, &e->G.ctr[r].problem, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int rb_sample(rb_engine *e, int32_t what, int32_t age, int32_t severity, int32_t n, int32_t *out) {
    CK(cudaSetDevice(e->cfg.device));
    if (what < 0 || what > 6 || age < 0 || age >= e->cfg.n_ages) { snprintf(g_err, sizeof g_err, "bad sample request"); return 1; }
    int32_t *d; CK(cudaMalloc(&d, sizeof(int32_t) * n));
    int epoch = e->h_sched[e->day > 0 ? e->day - 1 : 0].table_epoch;
    k_sample<<<(n + 255) / 256, 256, 0, e->stream>>>(e->G, what, age, severity, n, epoch, d); e->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out, d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

extern "C" int rb_read_agents(rb_engine *e, int32_t replica, rb_agent *out) {
    CK(cudaSetDevice(e->cfg.device));
    const Eng &G = e->G;
    if (replica < 0 || replica >= G.R) { snprintf(g_err, sizeof g_err, "bad replica"); return 1; }
    CK(cudaStreamSynchronize(e->stream));
    const int N = G.N; const size_t base = (size_t)replica * G.Npad;
    std::vector<uint32_t> hot(N), cold(N); std::vector<int32_t> inf(N); std::vector<int16_t> vd(N);
    CK(cudaMemcpy(hot.data(), G.hot + base, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost));
    {
        std::vector<AgentRec> rec(N);
        CK(cudaMemcpy(rec.data(), G.rec + base, sizeof(AgentRec) * N, cudaMemcpyDeviceToHost));
        for (int a = 0; a < N; a++) { cold[a] = rec[a].cold; inf[a] = rec[a].infector; vd[a] = rec[a].vacc_day; }
    }
    for (int a = 0; a < N; a++) {
        uint32_t h = hot[a]; rb_agent *o = &out[a];
        o->infector = inf[a]; o->n_infected = (int32_t)(cold[a] & 0xffffu);
        o->days_left = (int16_t)H_DL(h); o->day_of_illness = (int16_t)H_DOI(h); o->day_of_vaccination = vd[a];
        o->state = (uint8_t)H_STATE(h); o->severity = (uint8_t)H_SEV(h); o->variant = (uint8_t)H_VAR(h);
        o->flags = (uint8_t)(((h & H_DET) ? 1 : 0) | ((h & H_QUEUED) ? 2 : 0) | ((h & H_INCL) ? 4 : 0) | ((h & H_LIST) ? 8 : 0));
        o->ward_days = (uint8_t)((cold[a] >> 16) & 255u); o->icu_days = (uint8_t)((cold[a] >> 24) & 255u);
    }
    return 0;
}

extern "C" int rb_read_queue(rb_engine *e, int32_t replica, int32_t *out, int32_t cap, int32_t *n) {
    CK(cudaSetDevice(e->cfg.device));
    const Eng &G = e->G;
    CK(cudaStreamSynchronize(e->stream));
    RepCtr c; CK(cudaMemcpy(&c, &G.ctr[replica], sizeof c, cudaMemcpyDeviceToHost));
    *n = (int32_t)c.n_queue;
    int m = (int)c.n_queue < cap ? (int)c.n_queue : cap;
    if (m > 0) CK(cudaMemcpy(out, G.q_agent + ((size_t)replica * 2 + c.qsel) * G.cap_queue, sizeof(int32_t) * m, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int rb_read_available(rb_engine *e, int32_t replica, int32_t *o) {
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    RepCtr c; CK(cudaMemcpy(&c, &e->G.ctr[replica], sizeof c, cudaMemcpyDeviceToHost));
    o[0] = c.avail_beds; o[1] = c.avail_icu;
    return 0;
}

// ---------------------------------------------------------------- checkpoint / resume
// The reference keeps its state only in process memory (no checkpointing); a device-resident run
// of many replicas is worth saving.  The blob is the engine's whole mutable state between two rb_step calls.
struct StateHeader {
    uint64_t magic; int32_t version, n_agents, n_replicas, n_ages, row_len, max_days, day, sus_words;
    uint32_t cap_queue; uint32_t seed; int32_t rec_bytes, ctr_bytes;
};
#define STATE_MAGIC 0x3030324252414e49ull      // "INARB200"
struct StatePart { void *dev; size_t bytes; };
static std::vector<StatePart> state_parts(rb_engine *e) {
    const Eng &G = e->G;
    const size_t RN = (size_t)G.R * G.Npad, RW = (size_t)G.R * G.sus_words, RQ = (size_t)G.R * 2 * G.cap_queue;
    return {
        {G.hot, RN * sizeof(uint32_t)}, {G.rec, RN * sizeof(AgentRec)}, {G.sus, RW * sizeof(uint32_t)}, {G.act, RW * sizeof(uint32_t)},
        {G.ctr, (size_t)G.R * sizeof(RepCtr)}, {G.q_key, RQ * sizeof(unsigned long long)}, {G.q_agent, RQ * sizeof(int32_t)},
        {G.stats, (size_t)G.R * (G.max_days + 1) * G.row_len * sizeof(int32_t)},
    };
}
static StateHeader state_header(rb_engine *e) {
    const Eng &G = e->G;
    StateHeader h; memset(&h, 0, sizeof h);
    h.magic = STATE_MAGIC; h.version = 1; h.n_agents = G.N; h.n_replicas = G.R; h.n_ages = G.n_ages; h.row_len = G.row_len;
    h.max_days = G.max_days; h.day = e->day; h.sus_words = G.sus_words; h.cap_queue = G.cap_queue; h.seed = e->cfg.seed;
    h.rec_bytes = (int32_t)sizeof(AgentRec); h.ctr_bytes = (int32_t)sizeof(RepCtr);
    return h;
}

extern "C" int64_t rb_state_bytes(rb_engine *e) {
    size_t n = sizeof(StateHeader);
    for (const StatePart &p : state_parts(e)) n += p.bytes;
    return (int64_t)n;
}

extern "C" int rb_save_state(rb_engine *e, void *out, int64_t capacity) {
    CK(cudaSetDevice(e->cfg.device));
    if (capacity < rb_state_bytes(e)) { snprintf(g_err, sizeof g_err, "rb_save_state: buffer of %lld bytes, need %lld", (long long)capacity, (long long)rb_state_bytes(e)); return 1; }
    CK(cudaStreamSynchronize(e->stream));
    uint8_t *o = (uint8_t *)out;
    const StateHeader h = state_header(e);
    memcpy(o, &h, sizeof h); o += sizeof h;
    for (const StatePart &p : state_parts(e)) { CK(cudaMemcpy(o, p.dev, p.bytes, cudaMemcpyDeviceToHost)); o += p.bytes; }
    return 0;
}

extern "C" int rb_load_state(rb_engine *e, const void *in, int64_t n_bytes) {
    CK(cudaSetDevice(e->cfg.device));
    if (n_bytes != rb_state_bytes(e)) { snprintf(g_err, sizeof g_err, "rb_load_state: %lld bytes, this engine's state is %lld", (long long)n_bytes, (long long)rb_state_bytes(e)); return 1; }
    StateHeader h; memcpy(&h, in, sizeof h);
    StateHeader w = state_header(e); w.day = h.day; w.seed = h.seed;
    if (memcmp(&h, &w, sizeof h) != 0) { snprintf(g_err, sizeof g_err, "rb_load_state: the blob was saved by an engine of another shape (agents / replicas / max_days / build)"); return 1; }
    if (h.day < 0 || h.day > e->cfg.max_days) { snprintf(g_err, sizeof g_err, "rb_load_state: bad day"); return 1; }
    CK(cudaStreamSynchronize(e->stream));
    const uint8_t *o = (const uint8_t *)in + sizeof h;
    for (const StatePart &p : state_parts(e)) { CK(cudaMemcpy(p.dev, o, p.bytes, cudaMemcpyHostToDevice)); o += p.bytes; }
    e->day = h.day; e->cfg.seed = h.seed;
    return 0;
}
