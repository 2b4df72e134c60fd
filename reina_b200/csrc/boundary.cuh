// reina_b200 / csrc / boundary.cuh
// The day boundary, one CTA per replica: end of day d (beds / ICU first-come-first-served) fused with the start of day
// d+1 (stats row, interventions, imports, test queue + contact tracing, vaccination, sweep start).
#ifndef REINA_B200_BOUNDARY_CUH
#define REINA_B200_BOUNDARY_CUH
#include "state.cuh"

// ---------------------------------------------------------------- the team that runs one replica's day boundary
// Narrow: one CTA (ensembles: a GPU full of replicas, one CTA each).  Wide: `ncta` co-resident CTAs of a cooperative
// launch joined by a grid barrier (few replicas of a large population: a day's test queue, tracing attempts and
// capacity events run to 10^5 entries at 5 x 10^7 agents, far too many for one CTA).  The heavy phases stride over
// the whole team; the small sequential ones stay on the lead CTA.  A wide launch falls back to its lead CTA alone on
// days with little to do (RepCtr::wide_day, decided by k_resolve), so quiet days pay for no grid barrier.
struct Team {
    uint32_t tid, nth;          // thread index / thread count across the team
    uint32_t cta, ncta;
    bool lead;                  // this CTA runs the sequential phases
    unsigned int *bar;          // grid barrier word (RepCtr::wide_bar), zero at kernel entry and exit
    unsigned int gen;           // barriers passed
};
__device__ __forceinline__ Team team_of(RepCtr *c, uint32_t cta, uint32_t ncta) {
    Team T;
    T.cta = cta; T.ncta = ncta; T.lead = cta == 0; T.bar = &c->wide_bar; T.gen = 0;
    T.tid = cta * blockDim.x + threadIdx.x; T.nth = ncta * blockDim.x;
    return T;
}
__device__ __forceinline__ void team_sync(Team &T) {
    __syncthreads();
    if (T.ncta > 1) {
        T.gen++;
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(T.bar, 1u);
            const unsigned int target = T.gen * T.ncta;
            while (*(volatile unsigned int *)T.bar < target) { }
            __threadfence();
        }
        __syncthreads();
    }
}
// last statement of a team kernel: leaves the barrier word at zero for the next launch
__device__ __forceinline__ void team_finish(Team &T) {
    if (T.ncta > 1) { team_sync(T); if (T.tid == 0) *T.bar = 0u; }
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

// The same sort by a whole team.  Narrow teams use the single-CTA version; a wide team keeps the bucket counters in
// global scratch, builds the histogram and scatters with every thread, scans the counters on the lead CTA, and lets
// every thread order and write back whole buckets.  Falls back to the lead CTA's bitonic sort if scratch is too small.
__device__ void team_sort_pairs(Team &T, unsigned long long *keys, int32_t *vals, uint32_t n, uint32_t cap, Attempt *scratch, uint32_t scratch_cap,
                                BucketMap bm, unsigned long long *sk, int32_t *sv, int *warp_sums) {
    if (T.ncta == 1) {
        if (n > 1 && !block_bucket_sort(keys, vals, n, scratch, scratch_cap, bm, sv, warp_sums)) block_sort_pairs(keys, vals, n, cap, sk, sv);
        __syncthreads();
        return;
    }
    if (n <= 1) return;
    uint32_t B = SORT_BUCKETS;
    while (B < n && B < SORT_BUCKETS_MAX) B <<= 1;
    // the segmented scan below needs B / ncta to be a power of two: teams whose size is not one sort on the lead CTA
    if (2ull * n + B / 4 + WIDE_MAX_CTAS / 4 + 1 > scratch_cap || (T.ncta & (T.ncta - 1)) != 0 || B / T.ncta < 4) {
        if (T.lead) { if (!block_bucket_sort(keys, vals, n, scratch, scratch_cap, bm, sv, warp_sums)) block_sort_pairs(keys, vals, n, cap, sk, sv); }
        team_sync(T);
        return;
    }
    bm.B = B;
    int32_t *cnt = (int32_t *)(scratch + 2ull * n);
    for (uint32_t b = T.tid; b < B; b += T.nth) cnt[b] = 0;
    team_sync(T);
    for (uint32_t i = T.tid; i < n; i += T.nth) {
        const unsigned long long k = keys[i];
        const uint32_t b = min(bucket_of(bm, k), B - 1u);
        const uint32_t slot = (uint32_t)atomicAdd(&cnt[b], 1);
        scratch[i].key = k; scratch[i].cand = (uint32_t)vals[i]; scratch[i].parent = b | (slot << 16);
    }
    team_sync(T);
    // exclusive scan of the counters: every CTA scans its own contiguous B / ncta of them (4 consecutive counters per
    // thread and pass) and publishes its total; the offsets of the CTAs before it are added where the scan is used
    int32_t *ctot = cnt + B;                                 // [ncta] totals, right behind the counters
    const uint32_t seg = B / T.ncta, s0 = T.cta * seg;       // B and ncta are powers of two / divide evenly (checked by the caller)
    {
        int run = 0;
        for (uint32_t t0 = 0; t0 < seg; t0 += 4u * blockDim.x) {
            const uint32_t b = s0 + t0 + 4u * threadIdx.x;
            int4 v = make_int4(0, 0, 0, 0);
            if (t0 + 4u * threadIdx.x < seg) v = *reinterpret_cast<const int4 *>(cnt + b);
            const int mine = v.x + v.y + v.z + v.w;
            int total;
            int ex = block_scan_incl(mine, &total, warp_sums) - mine + run;
            if (t0 + 4u * threadIdx.x < seg) {
                int4 o; o.x = ex; o.y = ex + v.x; o.z = o.y + v.y; o.w = o.z + v.z;
                *reinterpret_cast<int4 *>(cnt + b) = o;
            }
            run += total;
        }
        if (threadIdx.x == 0) ctot[T.cta] = run;
    }
    team_sync(T);
    if (threadIdx.x < 32) {                                  // offsets of the segments: prefix over <= 64 CTA totals, kept in shared memory
        int acc = 0;
        for (uint32_t k0 = 0; k0 < T.ncta; k0 += 32) {
            const uint32_t k = k0 + threadIdx.x;
            int v = k < T.ncta ? ctot[k] : 0, incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += t; }
            if (k < T.ncta) sv[k] = acc + incl - v;
            acc += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    const int32_t *segoff = sv;
    uint32_t seg_shift = 0; while ((1u << seg_shift) < seg) seg_shift++;
    Attempt *out = scratch + n;
    for (uint32_t i = T.tid; i < n; i += T.nth) {
        const Attempt e = scratch[i];
        const uint32_t b = e.parent & 0xffffu;
        out[segoff[b >> seg_shift] + cnt[b] + (e.parent >> 16)] = e;
    }
    team_sync(T);
    for (uint32_t b = T.tid; b < B; b += T.nth) {          // insertion sort inside each bucket, written straight back
        const uint32_t lo = (uint32_t)(segoff[b >> seg_shift] + cnt[b]);
        const uint32_t hi = b + 1 < B ? (uint32_t)(segoff[(b + 1) >> seg_shift] + cnt[b + 1]) : n;
        for (uint32_t i = lo + 1; i < hi; i++) {
            const Attempt e = out[i];
            uint32_t j = i;
            while (j > lo && out[j - 1].key > e.key) { out[j] = out[j - 1]; j--; }
            out[j] = e;
        }
        for (uint32_t i = lo; i < hi; i++) { keys[i] = out[i].key; vals[i] = (int32_t)out[i].cand; }
    }
    team_sync(T);
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
#define IMP_CHUNK PRE_THREADS
__device__ __forceinline__ int32_t import_draw(const Eng &G, const RepCtr *c, size_t base, uint32_t ord, uint32_t t) {
    u32x4 x = philox(c->seed, ord, (uint32_t)c->day, PU_IMPORT | (t << 8), 0);
    float p = u01f(x.x);
    int k = G.n_import_classes - 1;
    for (int j = 0; j < G.n_import_classes; j++) if (p <= G.import_cum[j]) { k = j; break; }
    int32_t s = G.age_start[G.import_lo[k]], en = G.age_start[G.import_hi[k] + 1];
    int32_t pi = s + (int32_t)(x.y % (uint32_t)(en - s));
    return H_STATE(G.hot[base + pi]) == RB_SUSCEPTIBLE ? pi : -1;
}
__device__ void import_infections(const Eng &G, int r, RepCtr *c, int count, int variant, bool has_list, int *ordinal, int32_t *chosen /*[IMP_CHUNK]*/,
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
        // imported before today's sweep: the new entry goes to the list the sweep is about to read, flagged `fresh`
        if (tid < m && chosen[tid] >= 0) device_infect(G, r, c, chosen[tid], -1, 0u, variant, 0, true, (int)c->lsel, has_list);
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

#define TS(k) do { if (G.dbg == 9 && T.tid == 0) { long long now_ = clock64(); c->dbg_t[k] += now_ - c->dbg_last; c->dbg_last = now_; } } while (0)
// ---------------------------------------------------------------- day boundary (one team per replica)
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
    MP pre_bed, pre_icu;
};

__device__ void pre_body(const Eng &G, const int r, SmemSmall &S, Team &T) {
    unsigned long long *sk = S.sk; int32_t *sv = S.sv; int32_t *srow = S.u.srow; int *warp_sums = S.warp_sums; int *sh_i = S.sh_i;
    RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const int day = c->day;
    const rb_day_params *dp = &G.sched[day];
    const int tid = threadIdx.x;

    // queue bookkeeping read before anybody changes it
    const uint32_t cur = c->qsel, nxt = cur ^ 1u;
    unsigned long long *qk = G.q_key + ((size_t)r * 2 + cur) * G.cap_queue;
    int32_t *qa = G.q_agent + ((size_t)r * 2 + cur) * G.cap_queue;
    unsigned long long *nk = G.q_key + ((size_t)r * 2 + nxt) * G.cap_queue;
    int32_t *na = G.q_agent + ((size_t)r * 2 + nxt) * G.cap_queue;
    const uint32_t nq = c->n_queue;

    if (G.dbg == 9 && T.tid == 0) c->dbg_last = clock64();
    if (T.lead) {      // the small sequential phases: one CTA
        write_stats_row(G, r, c, srow);
        TS(0);

        // apply_intervention effects dated today (main.pyx:1880-1960), then Population.init_day (:1687-1699)
        if (tid == 0) {
            c->testing_mode = dp->testing_mode;
            if (dp->testing_mode == RB_ALL_WITH_SYMPTOMS_CT) c->ct_ever = 1u;
            c->p_detected_anyway = dp->p_detected_anyway;
            c->p_successful_tracing = dp->p_successful_tracing;
            c->beds += dp->beds_delta; c->avail_beds += dp->beds_delta;
            c->icu += dp->icu_delta; c->avail_icu += dp->icu_delta;
        }
        __syncthreads();
        int ordinal = 0;       // uniform across the CTA
        for (int i = 0; i < dp->n_imports; i++)      // in list order; each sees the testing mode in force at its turn (main.pyx:2012-2015)
            import_infections(G, r, c, dp->import_amount[i], dp->import_variant[i], (dp->import_traced >> i) & 1, &ordinal, sv, &sh_i[2]);
        __syncthreads();
        for (int i = tid; i < G.n_ages; i += blockDim.x) { c->counts[RB_A_NEW_INFECTIONS][i] = 0; c->counts[RB_A_DETECTED][i] = 0; }
        if (tid < RB_N_PLACES) c->daily_contacts[tid] = 0;
        if (tid < RB_MAX_VARIANTS) c->by_variant[tid] = 0;
        __syncthreads();
        if (tid == 0) { c->epoch = dp->table_epoch; c->total_infectors = 0; c->total_infections = 0; c->exposed_per_day = 0; }
        for (int v = 0; v < G.n_variants; v++)
            if (dp->trickle[v]) import_infections(G, r, c, dp->trickle[v], v, dp->testing_mode == RB_ALL_WITH_SYMPTOMS_CT, &ordinal, sv, &sh_i[2]);
        __syncthreads();
        TS(1);   // imports + init_day
        if (tid == 0) { c->ct_cases = (int32_t)nq; c->n_newq = 0; c->n_l0 = 0; c->n_l1 = 0; c->n_edges = 0; }
    }
    team_sync(T);

    // HealthcareSystem.iterate (main.pyx:514-558): drain yesterday's queue
    const bool ct = c->testing_mode == RB_ALL_WITH_SYMPTOMS_CT;
    if (ct && nq > 1) {                                            // queue order only matters for tracing
        BucketMap bm; bm.n_agents = (uint32_t)G.N; bm.n_prev = c->n_queue_prev; bm.kind = 1;
        team_sort_pairs(T, qk, qa, nq, G.cap_queue, G.succ + (size_t)r * G.cap_succ, G.cap_succ, bm, sk, sv, warp_sums);
    }
    TS(10);  // queue sort
    if (c->drained) {      // k_resolve already marked the queued agents detected: book the counts here, where the reference drains
        if (T.lead)
            for (int age = tid; age < G.n_ages; age += blockDim.x) {
                const int d = c->drain_det[age];
                if (d) { c->counts[RB_A_DETECTED][age] += d; c->counts[RB_A_ALL_DETECTED][age] += d; c->drain_det[age] = 0; }
            }
    } else
    for (uint32_t i = T.tid; i < nq; i += T.nth) {
        int32_t a = qa[i];
        uint32_t h = G.hot[base + a];
        if (h & H_DET) set_problem(c, RB_WRONG_STATE);   // person_detect, main.pyx:294-298
        G.hot[base + a] = (h & ~H_QUEUED) | H_DET;
        mark_detected(G, r, a);
        int age = age_of(G, a);
        count_add(c, RB_A_DETECTED, age, 1); count_add(c, RB_A_ALL_DETECTED, age, 1);
    }
    team_sync(T);
    if (T.tid == 0) c->drained = 0u;

    TS(2);   // queue drain
    if (ct && nq > 0) {
        // Depth-first contact tracing resolved in parallel.  Attempt key = (queue rank, level-0 slot, level-1 slot);
        // an attempt queues its candidate iff it is the smallest-key LIVE attempt on it that EXISTS; a level-1
        // attempt exists iff its tracer was itself queued by a level-0 attempt (main.pyx:498-499, 505-512).
        Attempt *l0 = G.succ + (size_t)r * G.cap_succ;
        Attempt *l1 = (Attempt *)(G.items + (size_t)r * G.cap_items);
        const uint32_t cap_l1 = G.cap_items / 2;
        const float ptr = c->p_successful_tracing;
        for (uint32_t i = T.tid; i < nq; i += T.nth) {
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
        team_sync(T);
        TS(5);   // tracing: level-0 attempts
        uint32_t n0 = min(c->n_l0, G.cap_succ);
        for (uint32_t j = T.tid; j < n0; j += T.nth) {
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
        team_sync(T);
        uint32_t n1 = min(c->n_l1, cap_l1);
        // kill edges: a level-1 attempt that precedes the level-0 winner of the same candidate
        uint32_t *esrc = (uint32_t *)(G.ev_key + (size_t)r * G.cap_events);
        uint32_t *edst = (uint32_t *)(G.ev_agent + (size_t)r * G.cap_events);
        const uint32_t cap_e = G.cap_events;
        for (uint32_t k = T.tid; k < n1; k += T.nth) {
            unsigned long long w = G.rec[base + (l1[k].cand)].winner;
            if (w != KEY_IDLE && l1[k].key < w) {
                uint32_t idx = atomicAdd(&c->n_edges, 1u);
                if (idx < cap_e) { esrc[idx] = l1[k].parent; edst[idx] = l1[k].cand; } else set_problem(c, RB_OTHER_FAILURE);
            }
        }
        team_sync(T);
        TS(6);   // tracing: level-1 attempts, kill edges
        uint32_t ne = min(c->n_edges, cap_e);
        if (ne > 0 && T.tid == 0) {
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
        if (ne > 0) team_sync(T);
        TS(7);   // tracing: cyclic dependencies (one thread)
        for (uint32_t k = T.tid; k < n1; k += T.nth) {
            Attempt e = l1[k];
            if (G.rec[base + (e.parent)].winner == (e.key & ~127ull)) atomicMin(&G.rec[base + (e.cand)].winner, e.key);
        }
        team_sync(T);
        for (uint32_t j = T.tid; j < n0 + n1; j += T.nth) {
            Attempt e = j < n0 ? l0[j] : l1[j - n0];
            if (G.rec[base + (e.cand)].winner != e.key) continue;
            if (j >= n0 && G.rec[base + (e.parent)].winner != (e.key & ~127ull)) continue;
            uint32_t idx = atomicAdd(&c->n_newq, 1u);
            if (idx < G.cap_queue) { nk[idx] = e.key; na[idx] = (int32_t)e.cand; } else set_problem(c, RB_OTHER_FAILURE);
            atomicOr(&G.hot[base + e.cand], H_QUEUED);
        }
        team_sync(T);
        for (uint32_t j = T.tid; j < n0 + n1; j += T.nth) { Attempt e = j < n0 ? l0[j] : l1[j - n0]; G.rec[base + (e.cand)].winner = KEY_IDLE; }
        team_sync(T);
    }

    TS(3);   // contact tracing
    // vaccinate_people (main.pyx:560-583): top-down walk of the age-sorted range; eligibility only ever turns
    // off (dead / vaccinated / detected), so a per-programme cursor below which the walk resumes is exact.
    if (T.lead)
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
    if (T.tid == 0) {
        u32x4 x = philox(c->seed, 0u, (uint32_t)day, PU_START, 0);     // _iterate_people, main.pyx:1988
        c->start = x.x % (uint32_t)G.N;
        c->n_items = 0; c->n_succ = 0; c->n_events = 0;
        c->n_queue_prev = nq;      // tomorrow's queue holds tracing keys whose rank field is < nq
        c->n_q_base = c->n_newq;
    }
    if (G.xbuf) {     // this rank's message header: the sweep and the contact kernel add to it from zero
        uint32_t *hw = (uint32_t *)xslot_of(G, G.rank, day).hdr;
        for (uint32_t i = T.tid; i < (uint32_t)(sizeof(RepCtr) / 4); i += T.nth) hw[i] = 0u;
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

__device__ void post_body(const Eng &G, const int r, SmemSmall &S, Team &T) {
    unsigned long long *sk = S.sk; int32_t *sv = S.sv; MP *s_bed = S.u.scan.bed, *s_icu = S.u.scan.icu;
    RepCtr *c = &G.ctr[r];
    const size_t base = (size_t)r * G.Npad;
    const int tid = threadIdx.x;
    const uint32_t n = min(c->n_events, G.cap_events);
    unsigned long long *ek = G.ev_key + (size_t)r * G.cap_events;
    int32_t *ea = G.ev_agent + (size_t)r * G.cap_events;
    const int day = c->day;
    const int beds0 = c->avail_beds, icu0 = c->avail_icu;
    const uint32_t n_newq = c->n_newq;
    const uint32_t old_l = c->lsel;
    if (G.dbg == 9 && T.tid == 0) c->dbg_last = clock64();
    // the lists today's sweep has read are emptied: tomorrow's sweep writes them (nobody reads these counters before then)
    for (uint32_t s = T.tid; s < G.n_seg; s += T.nth) *seg_count(G, r, old_l, s) = make_uint2(0u, 0u);
    if (n > 0) {
        BucketMap bm; bm.n_agents = (uint32_t)G.N; bm.n_prev = 0; bm.kind = 0;
        team_sort_pairs(T, ek, ea, n, G.cap_events, G.succ + (size_t)r * G.cap_succ, G.cap_succ, bm, sk, sv, S.warp_sums);
        TS(8);   // event sort
        // every thread of the team owns `per` consecutive events: compose them, scan the composed maps inside the CTA,
        // and (wide) chain the CTAs' totals through RepCtr::wide_mp
        const uint32_t per = (n + T.nth - 1) / T.nth;
        const uint32_t lo = min(n, T.tid * per), hi = min(n, lo + per);
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
        MP idm; idm.a = 0; idm.b = NEG_INF;
        if (T.ncta > 1) {
            if (tid == 0) {
                const MP tb = s_bed[blockDim.x - 1], ti = s_icu[blockDim.x - 1];
                c->wide_mp[T.cta][0] = tb.a; c->wide_mp[T.cta][1] = tb.b; c->wide_mp[T.cta][2] = ti.a; c->wide_mp[T.cta][3] = ti.b;
            }
            team_sync(T);
            if (tid == 0) {
                MP pb = idm, pi = idm;
                for (uint32_t k = 0; k < T.cta; k++) {
                    MP tb, ti; tb.a = c->wide_mp[k][0]; tb.b = c->wide_mp[k][1]; ti.a = c->wide_mp[k][2]; ti.b = c->wide_mp[k][3];
                    pb = mp_compose(pb, tb); pi = mp_compose(pi, ti);
                }
                S.pre_bed = pb; S.pre_icu = pi;
            }
            __syncthreads();
        } else if (tid == 0) { S.pre_bed = idm; S.pre_icu = idm; }
        __syncthreads();
        int beds = mp_apply(tid > 0 ? mp_compose(S.pre_bed, s_bed[tid - 1]) : S.pre_bed, beds0);
        int icu = mp_apply(tid > 0 ? mp_compose(S.pre_icu, s_icu[tid - 1]) : S.pre_icu, icu0);
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
        // the last thread of the team has composed everything before it; its running counters are the day's result
        if (T.tid == T.nth - 1) { c->avail_beds = beds; c->avail_icu = icu; }
    }
    team_sync(T);
    TS(9);   // capacity scan + outcomes
    if (T.tid == 0) {
        c->qsel ^= 1u;
        c->n_queue = min(n_newq, G.cap_queue);
        c->n_newq = 0;
        c->lsel = old_l ^ 1u;        // the lists today's sweep and k_resolve wrote become tomorrow's
        c->day = day + 1;            // main.pyx:2009
    }
}

// Narrow launch: grid = replicas.  Wide launch (cooperative): grid = (CTAs per replica, replicas); on a quiet day
// only the lead CTA stays.  k_pre decides from the queue it is about to drain (nothing in k_pre changes n_queue).
template <bool WIDE> __device__ __forceinline__ bool boundary_team(const Eng &G, int &r, Team &T, bool from_queue) {
    if (!WIDE) { r = blockIdx.x + G.r0; T = team_of(&G.ctr[r], 0, 1); return true; }
    r = blockIdx.y + G.r0;
    RepCtr *c = &G.ctr[r];
    const bool wide = from_queue ? c->n_queue >= (uint32_t)G.wide_min : c->wide_day != 0u;
    if (!wide && blockIdx.x != 0) return false;
    T = team_of(c, blockIdx.x, wide ? gridDim.x : 1);
    return true;
}
template <bool WIDE> __global__ void __launch_bounds__(PRE_THREADS) k_pre(Eng G) {
    __shared__ SmemSmall S;
    int r; Team T;
    if (!boundary_team<WIDE>(G, r, T, true)) return;
    pre_body(G, r, S, T);
    team_finish(T);
}
template <bool WIDE> __global__ void __launch_bounds__(PRE_THREADS) k_post(Eng G) {
    __shared__ SmemSmall S;
    int r; Team T;
    if (!boundary_team<WIDE>(G, r, T, false)) return;
    post_body(G, r, S, T);
    team_finish(T);
}
// end of day d (capacity scan) fused with the start of day d+1 (stats row, queue, tracing, ...): one launch less per day
template <bool WIDE> __global__ void __launch_bounds__(PRE_THREADS) k_between(Eng G) {
    __shared__ SmemSmall S;
    int r; Team T;
    if (!boundary_team<WIDE>(G, r, T, false)) return;
    post_body(G, r, S, T);
    team_sync(T);
    pre_body(G, r, S, T);
    team_finish(T);
}

#endif
