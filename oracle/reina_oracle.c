/*
 * reina_oracle.c -- CPU oracle for the per-day agent loop.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (reina_b200/) never does and fails loudly when its CUDA library is missing.
 *
 * What it is: a plain-C, single-threaded, SEQUENTIAL restatement of the reference algorithm
 * (cythonsim/main.pyx, cited per function as main.pyx:LINE), agent by agent in the reference's sweep
 * order, with first-come-first-served beds/ICU, first-infector-wins, depth-first contact tracing and the
 * reference's quirks (SURVEY.md section 8a notes 1-9).  The one deliberate change is the random stream:
 * the reference draws from a single sequential PCG64 (simrandom.pyx:13-55), which no parallel schedule
 * can reproduce, so every draw here comes from a counter-based Philox4x32-10 block keyed on
 * (seed; agent-or-ordinal, day, purpose|slot, iteration).  Because the draws no longer depend on the
 * order of evaluation, the CUDA engine -- which resolves the same order-dependent steps with
 * atomicMin on sweep positions, a sorted max-plus scan and a fixed-point contact-tracing resolve -- must
 * reproduce this oracle BIT-EXACTLY (every daily series and every agent field), and does in tests/.
 *
 * Pinning: the reference ships no golden vectors for this path ("parity unpinned").
 * The oracle is pinned STATISTICALLY against outputs of the unmodified reference engine run in the build
 * container (tests/golden/ref_ensemble_*.npz, made by tests/golden/make_golden.py from oracle/_ref):
 * ensemble means of every daily series within 3 standard errors, plus KS tests of the duration samplers
 * against numpy's gamma.
 *
 * Floating point: float ops are plain IEEE single (compile with -ffp-contract=off); log/exp are the
 * hand-rolled rb_logf/rb_expf below, restated identically in the CUDA engine so both sides round alike.
 * Agents are stored age-sorted (agent index = position in people_sorted_by_age, main.pyx:1431-1448); the
 * reference's shuffled `idx` order (np.random.shuffle, :1436) is replaced by a keyed Feistel permutation.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/reina_b200.h"

#define MAX_INFECTEES 64   /* main.pyx:128 */
#define MAX_CONTACTS 128   /* main.pyx:129 */

enum { PU_START = 1, PU_NCONTACT, PU_CONTACT, PU_SEVERITY, PU_INCUB, PU_ONSET, PU_SEEK, PU_NOBED,
       PU_TRACE, PU_IMPORT, PU_PERM, PU_SAMPLE, PU_CONTACT2, PU_INIT };
#define KEY1 0x5EEDB200u

typedef struct {
    int32_t infector, n_infected;
    int16_t days_left, day_of_illness, day_of_vaccination, day_of_infection;
    uint8_t state, severity, variant, detected, queued, included, has_list, age;
    uint8_t ward_days, icu_days;
    int32_t n_infectees;
    int32_t *infectees;
} Agent;

typedef struct {
    int32_t n_rows[RB_MAX_AGES];
    double cum_p[RB_MAX_AGES][RB_MAX_ROWS];
    int32_t start[RB_MAX_AGES][RB_MAX_ROWS], size[RB_MAX_AGES][RB_MAX_ROWS];
    uint8_t place[RB_MAX_AGES][RB_MAX_ROWS];
    float mask_p[RB_MAX_AGES][RB_MAX_ROWS];
    double nr_contacts[RB_MAX_AGES];
    double ncdf[RB_MAX_AGES][2][RB_NCDF];
    int set;
} Table;

typedef struct {
    uint32_t seed;
    Agent *agents;
    uint32_t fkey[4];
    int32_t *order;              /* absolute sweep slot -> agent */
    int32_t *perm;               /* agent -> absolute sweep slot */
    int32_t counts[RB_N_ATTRS][RB_MAX_AGES];
    int32_t daily_contacts[RB_N_PLACES], infected_by_variant[RB_MAX_VARIANTS];
    int32_t beds, icu, avail_beds, avail_icu;
    int32_t total_infectors, total_infections, exposed_per_day, ct_cases;
    int32_t problem;
    int32_t *queue, n_queue, cap_queue;
    int32_t *newq, n_newq, cap_newq;
    int32_t import_ordinal;
    int32_t list_on;             /* new cases get an infectee list (contact tracing in force at the moment of infection) */
    int32_t testing_mode;
    float p_detected_anyway, p_successful_tracing;
    int32_t epoch;
} Replica;

struct rb_engine {
    rb_config cfg;
    int32_t age_start[RB_MAX_AGES + 1];
    int32_t group_of_age[RB_MAX_AGES];
    rb_variant variants[RB_MAX_VARIANTS];
    int32_t import_lo[RB_MAX_IMPORT_CLASSES], import_hi[RB_MAX_IMPORT_CLASSES];
    float import_cum[RB_MAX_IMPORT_CLASSES];
    Table *tables; int n_tables;
    rb_day_params *sched; int n_sched;
    Replica *rep;
    int32_t day;
    int32_t *stats;   /* [replica][max_days+1][row_len] */
    int32_t row_len;
    int feistel_half;
    float last_ms;
    int32_t ipc[7]; int has_ipc;   /* initial population condition, re-applied by ro_reset */
};

static char g_err[256];
const char *ro_last_error(void) { return g_err; }

/* ------------------------------------------------------------------ RNG (replaces simrandom.pyx:13-55) */
static inline void philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline double u01d(uint32_t hi, uint32_t lo) {   /* next_double: 53 bits in [0,1) */
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
static inline float u01f(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }       /* [0,1) */
static inline float u01f_open(uint32_t x) { return (float)((x >> 8) + 1u) * (1.0f / 16777216.0f); } /* (0,1] */

static inline float rb_logf(float x) {
    uint32_t b; memcpy(&b, &x, 4);
    int e = (int)(b >> 23) - 127;
    b = (b & 0x007fffffu) | 0x3f800000u;
    float m; memcpy(&m, &b, 4);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float s = (m - 1.0f) / (m + 1.0f);
    float s2 = s * s;
    float p = 0.111111111f;
    p = p * s2 + 0.142857143f;
    p = p * s2 + 0.2f;
    p = p * s2 + 0.333333333f;
    p = p * s2 + 1.0f;
    return (float)e * 0.693147181f + (2.0f * s) * p;
}
static inline float rb_expf(float y) {
    float k = floorf(y * 1.44269504f + 0.5f);
    float r = y - k * 0.693359375f;
    r = r - k * -2.12194440e-4f;
    float p = 1.0f / 720.0f;
    p = p * r + 1.0f / 120.0f;
    p = p * r + 1.0f / 24.0f;
    p = p * r + 1.0f / 6.0f;
    p = p * r + 0.5f;
    p = p * r + 1.0f;
    p = p * r + 1.0f;
    int ki = (int)k;
    if (ki < -126) return 0.0f;
    if (ki > 127) ki = 127;
    uint32_t b = (uint32_t)(ki + 127) << 23;
    float sc; memcpy(&sc, &b, 4);
    return p * sc;
}
/* Marsaglia polar normal from one Philox block; returns 0 when the pair is rejected. */
static inline int polar_normal(const uint32_t x[4], float *z) {
    float u = 2.0f * u01f(x[0]) - 1.0f, v = 2.0f * u01f(x[1]) - 1.0f;
    float s = u * u + v * v;
    if (s >= 1.0f || s == 0.0f) return 0;
    *z = u * sqrtf(-2.0f * rb_logf(s) / s);
    return 1;
}
/* random_gamma_f(kappa, theta), kappa > 1: Marsaglia-Tsang (numpy legacy distributions), one Philox
 * block per trial: words 0,1 -> normal, word 2 -> acceptance uniform. */
static float gamma_f(uint32_t seed, uint32_t c0, uint32_t c1, uint32_t purpose, float kappa, float theta) {
    float d = kappa - 0.333333333f;
    float c = 1.0f / sqrtf(9.0f * d);
    for (uint32_t it = 0;; it++) {
        uint32_t x[4];
        philox(seed, KEY1, c0, c1, purpose, it, x);
        float z;
        if (!polar_normal(x, &z)) continue;
        float v = 1.0f + c * z;
        if (v <= 0.0f) continue;
        v = v * v * v;
        float u = u01f_open(x[2]);
        float z2 = z * z;
        if (u < 1.0f - 0.0331f * (z2 * z2)) return (d * v) * theta;
        if (rb_logf(u) < 0.5f * z2 + d * ((1.0f - v) + rb_logf(v))) return (d * v) * theta;
    }
}
static inline int chance(double u, float p) {   /* RandomPool.chance, simrandom.pyx:32-39 */
    if (p == 1.0f) return 1;
    if (p == 0.0f) return 0;
    return u < (double)p;
}
static inline int round_to_int(float f) { return (int)(f + 0.5f); }   /* main.pyx:773-774 */
static inline int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* ------------------------------------------------------------------ sweep order */
static inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
static uint32_t feistel(uint32_t a, uint32_t n, int half, const uint32_t k[4]) {
    uint32_t mask = (1u << half) - 1u, x = a;
    do {
        uint32_t L = x >> half, R = x & mask;
        for (int r = 0; r < 4; r++) { uint32_t t = L ^ (mix32(R ^ k[r]) & mask); L = R; R = t; }
        x = (L << half) | R;
    } while (x >= n);
    return x;
}

/* ------------------------------------------------------------------ model pieces */
static inline int age_of(const rb_engine *e, int32_t a) {
    int lo = 0, hi = e->cfg.n_ages;   /* largest age with age_start[age] <= a */
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (e->age_start[mid] <= a) lo = mid; else hi = mid; }
    return lo;
}
static inline int is_infected(const Agent *p) { return p->state >= RB_INCUBATION && p->state <= RB_IN_ICU; }

/* Disease.get_source_infectiousness, main.pyx:895-906 */
static inline float source_infectiousness(const rb_engine *e, const Agent *p) {
    int day;
    if (p->state == RB_INCUBATION) day = -p->days_left;
    else if (p->state == RB_ILLNESS) day = p->day_of_illness;
    else return 0.0f;
    if (day < -10 || day > 10) return 0.0f;
    return e->variants[p->variant].iot[day + 10];
}

/* Disease.get_symptom_severity, main.pyx:1042-1091 (both FATAL branches test the same inequality and the
 * first marks DEATH_OUTSIDE_HOSPITAL, so every FATAL case dies outside hospital: place of death is implied). */
static int symptom_severity(const rb_variant *v, int age, float val, int vaccinated_effective) {
    float vmod = 1.0f;
    if (vaccinated_effective) vmod = vmod * 0.1f;
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

static void counts_add(Replica *r, int attr, int age, int d) { r->counts[attr][age] += d; }

/* person_infect, main.pyx:209-235 (+ Population.infect :1576-1582).  Severity and incubation are drawn
 * while person.variant_idx is still 0, i.e. with wild-type parameters, exactly as the reference does. */
static void person_infect(rb_engine *e, Replica *r, int32_t ti, int32_t src, int variant, int slot_unused) {
    (void)slot_unused;
    Agent *t = &r->agents[ti];
    const rb_variant *v0 = &e->variants[0];
    uint32_t x[4];
    t->state = RB_INCUBATION;
    philox(r->seed, KEY1, (uint32_t)ti, (uint32_t)e->day, PU_SEVERITY, 0, x);
    int vacc_eff = t->day_of_vaccination >= 0 && (e->day - t->day_of_vaccination) > 14;
    t->severity = (uint8_t)symptom_severity(v0, t->age, u01f(x[0]), vacc_eff);
    t->days_left = (int16_t)clamp255(round_to_int(gamma_f(r->seed, (uint32_t)ti, (uint32_t)e->day, PU_INCUB,
                                                           v0->incubation_kappa, v0->incubation_theta)));
    t->day_of_infection = (int16_t)e->day;
    if (src >= 0) {
        Agent *s = &r->agents[src];
        t->infector = src;
        if (s->has_list) {
            if (s->n_infectees >= MAX_INFECTEES) { r->problem = RB_TOO_MANY_INFECTEES; }
            else s->infectees[s->n_infectees++] = ti;
        }
        variant = s->variant;
    }
    t->variant = (uint8_t)variant;
    if (r->list_on) {     /* hc.testing_mode == ALL_WITH_SYMPTOMS_CT at the moment of infection, main.pyx:227-233 */
        t->has_list = 1;
        t->infectees = (int32_t *)malloc(sizeof(int32_t) * MAX_INFECTEES);
        t->n_infectees = 0;
    }
    counts_add(r, RB_A_SUSCEPTIBLE, t->age, -1);
    counts_add(r, RB_A_INFECTED, t->age, 1);
    counts_add(r, RB_A_ALL_INFECTED, t->age, 1);
    counts_add(r, RB_A_NEW_INFECTIONS, t->age, 1);
    r->infected_by_variant[variant] += 1;
}

static void person_become_removed(Agent *p) {   /* main.pyx:301-307 */
    if (p->infectees) { free(p->infectees); p->infectees = NULL; }
    p->has_list = 0;
}
static void person_recover(Replica *r, Agent *p) {   /* main.pyx:315-318, Population.recover :1585-1588 */
    p->state = RB_RECOVERED; person_become_removed(p);
    counts_add(r, RB_A_INFECTED, p->age, -1); counts_add(r, RB_A_RECOVERED, p->age, 1);
}
static void person_die(Replica *r, Agent *p) {       /* main.pyx:370-374, Population.die :1618-1623 */
    p->state = RB_DEAD; person_become_removed(p);
    counts_add(r, RB_A_INFECTED, p->age, -1); counts_add(r, RB_A_DEAD, p->age, 1);
    if (p->severity == RB_FATAL) counts_add(r, RB_A_NON_HOSPITAL_DEATHS, p->age, 1);
}
static void person_detect(Replica *r, Agent *p) {    /* main.pyx:294-298, Population.detect :1591-1594 */
    if (p->detected) r->problem = RB_WRONG_STATE;
    p->detected = 1;
    counts_add(r, RB_A_DETECTED, p->age, 1); counts_add(r, RB_A_ALL_DETECTED, p->age, 1);
}
/* Disease.dies_in_hospital, main.pyx:957-974 */
static int dies_in_hospital(rb_engine *e, Replica *r, int32_t ai, int care_available) {
    Agent *p = &r->agents[ai];
    const rb_variant *v = &e->variants[p->variant];
    float ch = 0.0f;
    if (p->severity == RB_FATAL) return 1;
    if (p->severity == RB_CRITICAL) { if (care_available) return 0; ch = v->p_icu_death_no_beds; }
    else if (p->severity == RB_SEVERE) { if (care_available) return 0; ch = v->p_hospital_death_no_beds; }
    uint32_t x[4];
    philox(r->seed, KEY1, (uint32_t)ai, (uint32_t)e->day, PU_NOBED, 0, x);
    return chance(u01d(x[0], x[1]), ch);
}

static void push(int32_t **buf, int32_t *n, int32_t *cap, int32_t v) {
    if (*n == *cap) { *cap = *cap ? *cap * 2 : 1024; *buf = (int32_t *)realloc(*buf, sizeof(int32_t) * (size_t)*cap); }
    (*buf)[(*n)++] = v;
}

/* HealthcareSystem.queue_for_testing, main.pyx:474-488; the draw is keyed (tracer, day, candidate). */
static int queue_for_testing(rb_engine *e, Replica *r, int32_t ci, int32_t tracer, float p_success) {
    Agent *c = &r->agents[ci];
    if (c->state == RB_DEAD || c->detected || c->queued) return 0;
    if (tracer >= 0) {
        uint32_t x[4];
        philox(r->seed, KEY1, (uint32_t)tracer, (uint32_t)e->day, PU_TRACE, (uint32_t)ci, x);
        if (!chance(u01d(x[0], x[1]), p_success)) return 0;
    }
    c->queued = 1;
    push(&r->newq, &r->n_newq, &r->cap_newq, ci);
    return 1;
}
/* HealthcareSystem.perform_contact_tracing, main.pyx:495-512 */
static void contact_tracing(rb_engine *e, Replica *r, int32_t pi, int level) {
    if (level > 1) return;
    Agent *p = &r->agents[pi];
    if (p->infector >= 0)
        if (queue_for_testing(e, r, p->infector, pi, r->p_successful_tracing))
            contact_tracing(e, r, p->infector, level + 1);
    if (p->infectees)
        for (int i = 0; i < p->n_infectees; i++) {
            int32_t ci = p->infectees[i];
            if (queue_for_testing(e, r, ci, pi, r->p_successful_tracing)) contact_tracing(e, r, ci, level + 1);
        }
}
/* HealthcareSystem.seek_testing, main.pyx:595-615 */
static void seek_testing(rb_engine *e, Replica *r, int32_t ai) {
    Agent *p = &r->agents[ai];
    int q = 0;
    if (r->testing_mode == RB_ALL_WITH_SYMPTOMS || r->testing_mode == RB_ALL_WITH_SYMPTOMS_CT) q = 1;
    else if (r->testing_mode == RB_ONLY_SEVERE_SYMPTOMS) {
        if (p->severity >= RB_SEVERE) q = 1;
        else {
            uint32_t x[4];
            philox(r->seed, KEY1, (uint32_t)ai, (uint32_t)e->day, PU_SEEK, 0, x);
            if (chance(u01d(x[0], x[1]), r->p_detected_anyway)) q = 1;
        }
    }
    if (q) queue_for_testing(e, r, ai, -1, 1.0f);
}

/* Population.get_import_infection_person + infect_people, main.pyx:1632-1665 */
static void infect_people(rb_engine *e, Replica *r, int count, int variant) {
    for (int i = 0; i < count; i++) {
        uint32_t ordinal = (uint32_t)r->import_ordinal++;
        int32_t found = -1;
        for (uint32_t t = 0; t < 10; t++) {
            uint32_t x[4];
            philox(r->seed, KEY1, ordinal, (uint32_t)e->day, PU_IMPORT | (t << 8), 0, x);
            float p = u01f(x[0]);
            int k = e->cfg.n_import_classes - 1;
            for (int j = 0; j < e->cfg.n_import_classes; j++) if (p <= e->import_cum[j]) { k = j; break; }
            int32_t s = e->age_start[e->import_lo[k]], en = e->age_start[e->import_hi[k] + 1];
            int32_t pi = s + (int32_t)(x[1] % (uint32_t)(en - s));
            if (r->agents[pi].state == RB_SUSCEPTIBLE) { found = pi; break; }
        }
        if (found >= 0) person_infect(e, r, found, -1, variant, 0);
    }
}

/* HealthcareSystem.vaccinate_people, main.pyx:560-583 (+ person_vaccinate :377-392) */
static void vaccinate_people(rb_engine *e, Replica *r, int nr, int min_age, int max_age) {
    int32_t s = e->age_start[min_age], en = e->age_start[max_age + 1];
    if (nr > en - s) nr = en - s;
    int done = 0;
    for (int32_t i = en - 1; done < nr && i >= s; i--) {
        Agent *p = &r->agents[i];
        if (p->state == RB_DEAD || p->day_of_vaccination >= 0 || p->detected) continue;
        p->day_of_vaccination = (int16_t)e->day;
        counts_add(r, RB_A_VACCINATED, p->age, 1);
        done++;
    }
}

/* person_hospitalize, main.pyx:321-338 (+ HealthcareSystem.hospitalize :617-621) */
static void person_hospitalize(rb_engine *e, Replica *r, int32_t ai) {
    Agent *p = &r->agents[ai];
    if (!p->detected) person_detect(r, p);
    if (r->avail_beds == 0) {
        if (dies_in_hospital(e, r, ai, 0)) person_die(r, p); else person_recover(r, p);
        return;
    }
    r->avail_beds -= 1;
    p->days_left = p->ward_days;
    p->state = RB_HOSPITALIZED;
    counts_add(r, RB_A_IN_WARD, p->age, 1);
}
/* person_transfer_to_icu, main.pyx:341-351 (+ to_icu :643-648): the ward bed is always freed first. */
static void person_transfer_to_icu(rb_engine *e, Replica *r, int32_t ai) {
    Agent *p = &r->agents[ai];
    r->avail_beds += 1;
    if (r->avail_icu == 0) {
        if (dies_in_hospital(e, r, ai, 0)) {
            counts_add(r, RB_A_IN_WARD, p->age, -1);
            person_die(r, p);
            return;
        }
    } else r->avail_icu -= 1;
    p->days_left = p->icu_days;
    counts_add(r, RB_A_IN_WARD, p->age, -1); counts_add(r, RB_A_IN_ICU, p->age, 1); counts_add(r, RB_A_CUM_ICU, p->age, 1);
    p->state = RB_IN_ICU;
}
/* person_release_from_hospital, main.pyx:354-367 */
static void person_release(rb_engine *e, Replica *r, int32_t ai) {
    Agent *p = &r->agents[ai];
    int death = dies_in_hospital(e, r, ai, 1);
    if (p->state == RB_IN_ICU) { counts_add(r, RB_A_IN_ICU, p->age, -1); r->avail_icu += 1; }
    else { counts_add(r, RB_A_IN_WARD, p->age, -1); r->avail_beds += 1; }
    if (death) person_die(r, p); else person_recover(r, p);
}

/* person_become_ill, main.pyx:284-291 + the duration helpers :989-1039 (durations are fixed here, from the
 * same onset-to-removed draw the reference stores in days_from_onset_to_removed). */
static void person_become_ill(rb_engine *e, Replica *r, int32_t ai) {
    Agent *p = &r->agents[ai];
    const rb_variant *v = &e->variants[p->variant];
    p->state = RB_ILLNESS;
    float T = (p->severity == RB_FATAL)
        ? gamma_f(r->seed, (uint32_t)ai, (uint32_t)e->day, PU_ONSET, v->onset_death_kappa, v->onset_death_theta)
        : gamma_f(r->seed, (uint32_t)ai, (uint32_t)e->day, PU_ONSET, v->onset_recovery_kappa, v->onset_recovery_theta);
    float f = T;
    if (p->severity != RB_ASYMPTOMATIC && p->severity != RB_MILD) f = f * v->ratio_before_hospitalisation;
    p->days_left = (int16_t)clamp255(round_to_int(f));
    float w = 0.0f, u = 0.0f;
    if (p->severity == RB_SEVERE) w = T * (1.0f - v->ratio_before_hospitalisation);
    else if (p->severity == RB_CRITICAL || p->severity == RB_FATAL) {
        w = T * v->ratio_in_ward;
        u = ((1.0f - v->ratio_in_ward) - v->ratio_before_hospitalisation) * T;
    }
    p->ward_days = (uint8_t)clamp255(round_to_int(w));
    p->icu_days = (uint8_t)clamp255(round_to_int(u));
    if (p->severity != RB_ASYMPTOMATIC && !p->detected) seek_testing(e, r, ai);
}

/* person_expose_others, main.pyx:247-281 with get_exposed_people :936-955, get_contacts :1539-1573,
 * get_nr_contacts :1308-1320, get_one_contact :1290-1304, did_infect :908-934. */
static int person_expose_others(rb_engine *e, Replica *r, int32_t ai) {
    Agent *p = &r->agents[ai];
    if (p->detected) return 0;
    float si = source_infectiousness(e, p);
    if (si == 0.0f) return 0;
    /* get_nr_contacts, main.pyx:1308-1320: one uniform against the tabulated distribution of n */
    int cls = 0, limit = 100;
    if (p->state == RB_ILLNESS && p->severity != RB_ASYMPTOMATIC) { cls = 1; limit = 5; }
    const Table *tb = &e->tables[r->epoch];
    int n;
    {
        uint32_t x[4];
        philox(r->seed, KEY1, (uint32_t)ai, (uint32_t)e->day, PU_NCONTACT, 0, x);
        double u = u01d(x[0], x[1]);
        const double *cdf = tb->ncdf[p->age][cls];
        int lo = 0, hi = limit;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (u < cdf[mid]) hi = mid; else lo = mid + 1; }
        n = lo;
    }
    const rb_variant *v = &e->variants[p->variant];
    if (p->severity == RB_ASYMPTOMATIC) si = si * v->p_asymptomatic_infection;
    int nrows = tb->n_rows[p->age];
    /* Thinning (distribution-preserving restatement of did_infect, main.pyx:908-934): no target can be infected with
     * probability above p_upper = si * max susceptibility * multiplier, so a contact first passes a coarse 8-bit
     * filter with probability q = k/256 >= p_upper, and only then is its person drawn and the transmission decided
     * with probability p/q.  Four contacts share one Philox block: 24 bits pick the row, 8 bits feed the filter. */
    float p_upper = (si * v->reserved[0]) * v->infectiousness_multiplier;
    int kq = (int)(p_upper * 256.0f) + 1;
    if (kq > 256) kq = 256;
    for (int slot = 0; slot < n; slot++) {
        uint32_t x[4];
        philox(r->seed, KEY1, (uint32_t)ai, (uint32_t)e->day, PU_CONTACT | ((uint32_t)(slot >> 2) << 8), 0, x);
        uint32_t word = x[slot & 3];
        double u = (double)(word >> 8) * (1.0 / 16777216.0);
        int row = nrows - 1;   /* the reference fails with CONTACT_PROBABILITY_FAILURE here (p ~ 1e-15) */
        for (int i = 0; i < nrows; i++) if (u < tb->cum_p[p->age][i]) { row = i; break; }
        r->daily_contacts[tb->place[p->age][row]] += 1;
        if ((int)(word & 255u) >= kq) continue;
        philox(r->seed, KEY1, (uint32_t)ai, (uint32_t)e->day, PU_CONTACT2 | ((uint32_t)slot << 8), 0, x);
        int32_t ti = tb->start[p->age][row] + (int32_t)(x[0] % (uint32_t)tb->size[p->age][row]);
        Agent *t = &r->agents[ti];
        if (t->state != RB_SUSCEPTIBLE) continue;       /* person_expose, main.pyx:238-244 */
        float pr = (si * v->tab[RB_T_SUSCEPTIBILITY][t->age]) * v->infectiousness_multiplier;
        if (!(((double)x[1] * (1.0 / 4294967296.0)) * (double)kq < (double)pr * 256.0)) continue;
        float mp = tb->mask_p[p->age][row];
        if (mp != 0.0f) {
            float a = mp * v->p_mask_protects_others, b = mp * v->p_mask_protects_wearer;
            float pm = (a + b) - a * b;
            if (chance((double)x[2] * (1.0 / 4294967296.0), pm)) continue;
        }
        person_infect(e, r, ti, ai, -1, slot);
        if (p->has_list && p->n_infected >= MAX_INFECTEES) { r->problem = RB_TOO_MANY_INFECTEES; break; }
        p->n_infected += 1;
    }
    return n;
}

/* person_advance, main.pyx:395-438 */
static int person_advance(rb_engine *e, Replica *r, int32_t ai) {
    Agent *p = &r->agents[ai];
    int exposed = 0;
    switch (p->state) {
    case RB_INCUBATION:
        if (p->day_of_infection == e->day) return 0;
        exposed = person_expose_others(e, r, ai);
        if (p->days_left > 0) p->days_left -= 1;
        if (p->days_left == 0) person_become_ill(e, r, ai);
        break;
    case RB_ILLNESS:
        exposed = person_expose_others(e, r, ai);
        if (p->day_of_illness < 31) p->day_of_illness += 1;
        if (p->days_left > 0) p->days_left -= 1;
        if (p->days_left == 0) {
            if (p->severity == RB_FATAL) person_die(r, p);
            else if (p->severity >= RB_SEVERE) person_hospitalize(e, r, ai);
            else person_recover(r, p);
        }
        break;
    case RB_HOSPITALIZED:
        if (p->days_left > 0) p->days_left -= 1;
        if (p->days_left == 0) {
            if (p->severity == RB_CRITICAL || p->severity == RB_FATAL) person_transfer_to_icu(e, r, ai);
            else person_release(e, r, ai);
        }
        break;
    case RB_IN_ICU:
        if (p->days_left > 0) p->days_left -= 1;
        if (p->days_left == 0) person_release(e, r, ai);
        break;
    default: break;
    }
    return exposed;
}

static void snapshot(rb_engine *e, int ri) {   /* Context.generate_state, main.pyx:1813-1857 */
    Replica *r = &e->rep[ri];
    int32_t *row = e->stats + ((size_t)ri * (e->cfg.max_days + 1) + e->day) * e->row_len;
    memset(row, 0, sizeof(int32_t) * e->row_len);
    for (int a = 0; a < RB_N_ATTRS; a++)
        for (int age = 0; age < e->cfg.n_ages; age++) row[a * e->cfg.n_groups + e->group_of_age[age]] += r->counts[a][age];
    int32_t *s = row + RB_N_ATTRS * e->cfg.n_groups;
    s[RB_S_AVAILABLE_ICU] = r->avail_icu; s[RB_S_AVAILABLE_BEDS] = r->avail_beds;
    s[RB_S_TOTAL_ICU] = r->icu; s[RB_S_TOTAL_BEDS] = r->beds;
    s[RB_S_TOTAL_INFECTIONS] = r->total_infections; s[RB_S_TOTAL_INFECTORS] = r->total_infectors;
    s[RB_S_EXPOSED_PER_DAY] = r->exposed_per_day; s[RB_S_CT_CASES_PER_DAY] = r->ct_cases;
    s[RB_S_TABLE_EPOCH] = r->epoch; s[RB_S_DAY] = e->day;
    for (int i = 0; i < RB_N_PLACES; i++) s[RB_S_CONTACTS0 + i] = r->daily_contacts[i];
    for (int i = 0; i < RB_MAX_VARIANTS; i++) s[RB_S_VARIANT0 + i] = r->infected_by_variant[i];
}

/* Context.iterate + _iterate, main.pyx:1994-2018 */
static void iterate_replica(rb_engine *e, int ri) {
    Replica *r = &e->rep[ri];
    const rb_day_params *dp = &e->sched[e->day];
    const int32_t N = e->cfg.n_agents;
    /* apply_intervention for interventions dated today, main.pyx:1880-1960 */
    r->testing_mode = dp->testing_mode;
    r->p_detected_anyway = dp->p_detected_anyway;
    r->p_successful_tracing = dp->p_successful_tracing;
    r->beds += dp->beds_delta; r->avail_beds += dp->beds_delta;
    r->icu += dp->icu_delta; r->avail_icu += dp->icu_delta;
    r->import_ordinal = 0;
    /* interventions of one date are applied in list order (main.pyx:2012-2015): an import sees the testing mode that was
     * in force when its turn came, which the host recorded per event */
    for (int i = 0; i < dp->n_imports; i++) {
        r->list_on = (dp->import_traced >> i) & 1;
        infect_people(e, r, dp->import_amount[i], dp->import_variant[i]);
    }
    r->list_on = r->testing_mode == RB_ALL_WITH_SYMPTOMS_CT;
    /* Population.init_day, main.pyx:1687-1699 */
    memset(r->daily_contacts, 0, sizeof r->daily_contacts);
    memset(r->counts[RB_A_NEW_INFECTIONS], 0, sizeof r->counts[0]);
    memset(r->counts[RB_A_DETECTED], 0, sizeof r->counts[0]);
    memset(r->infected_by_variant, 0, sizeof r->infected_by_variant);
    r->epoch = dp->table_epoch;
    for (int v = 0; v < e->cfg.n_variants; v++) if (dp->trickle[v]) infect_people(e, r, dp->trickle[v], v);
    r->total_infectors = r->total_infections = r->exposed_per_day = 0;
    /* HealthcareSystem.iterate, main.pyx:514-558 */
    {
        /* r->queue = yesterday's list, r->newq = today's (empty) list */
        int32_t nq = r->n_queue;
        r->n_newq = 0;
        r->ct_cases = nq;
        for (int32_t i = 0; i < nq; i++) {
            int32_t ai = r->queue[i];
            Agent *p = &r->agents[ai];
            p->queued = 0;
            person_detect(r, p);
            if (r->testing_mode == RB_ALL_WITH_SYMPTOMS_CT) contact_tracing(e, r, ai, 0);
        }
        for (int i = 0; i < dp->n_vacc; i++)
            if (dp->vacc_nr[i]) vaccinate_people(e, r, dp->vacc_nr[i], dp->vacc_min_age[i], dp->vacc_max_age[i]);
    }
    /* _iterate_people + _process_person, main.pyx:1968-1992 */
    uint32_t x[4];
    philox(r->seed, KEY1, 0, (uint32_t)e->day, PU_START, 0, x);
    int32_t start = (int32_t)(x[0] % (uint32_t)N);
    for (int32_t i = 0; i < N; i++) {
        int32_t slot = start + i; if (slot >= N) slot -= N;
        int32_t ai = r->order[slot];
        Agent *p = &r->agents[ai];
        if ((p->state == RB_RECOVERED || p->state == RB_DEAD) && !p->included) {
            r->total_infectors += 1; r->total_infections += p->n_infected; p->included = 1;
        }
        if (!is_infected(p)) continue;
        r->exposed_per_day += person_advance(e, r, ai);
    }
    /* today's new queue becomes the queue drained tomorrow */
    { int32_t *t = r->queue; r->queue = r->newq; r->newq = t;
      int32_t c = r->cap_queue; r->cap_queue = r->cap_newq; r->cap_newq = c;
      r->n_queue = r->n_newq; r->n_newq = 0; }
}

/* ------------------------------------------------------------------ C-ABI (ro_ mirrors rb_) */
static void init_replica(rb_engine *e, int ri, uint32_t seed);
int ro_create(const rb_config *cfg, const int32_t *age_counts, const int32_t *group_of_age,
              const rb_variant *variants, const int32_t *import_lo, const int32_t *import_hi,
              const float *import_cum, rb_engine **out) {
    if (cfg->n_ages > RB_MAX_AGES || cfg->n_variants > RB_MAX_VARIANTS || cfg->n_import_classes > RB_MAX_IMPORT_CLASSES) {
        snprintf(g_err, sizeof g_err, "config exceeds compiled limits"); return 1;
    }
    rb_engine *e = (rb_engine *)calloc(1, sizeof *e);
    e->cfg = *cfg;
    int64_t tot = 0;
    for (int a = 0; a < cfg->n_ages; a++) { e->age_start[a] = (int32_t)tot; tot += age_counts[a]; e->group_of_age[a] = group_of_age[a]; }
    e->age_start[cfg->n_ages] = (int32_t)tot;
    if (tot != cfg->n_agents) { snprintf(g_err, sizeof g_err, "age_counts sum %lld != n_agents %d", (long long)tot, cfg->n_agents); free(e); return 1; }
    memcpy(e->variants, variants, sizeof(rb_variant) * cfg->n_variants);
    for (int i = 0; i < cfg->n_import_classes; i++) { e->import_lo[i] = import_lo[i]; e->import_hi[i] = import_hi[i]; e->import_cum[i] = import_cum[i]; }
    /* an import class that can be drawn (positive weight, or the last one: the fallback) must hold somebody */
    for (int i = 0; i < cfg->n_import_classes; i++) {
        float w = import_cum[i] - (i ? import_cum[i - 1] : 0.0f);
        if (import_lo[i] < 0 || import_hi[i] >= cfg->n_ages || import_lo[i] > import_hi[i]) { snprintf(g_err, sizeof g_err, "import class %d: bad age band", i); free(e); return 1; }
        if ((w > 0.0f || i == cfg->n_import_classes - 1) && e->age_start[import_hi[i] + 1] - e->age_start[import_lo[i]] <= 0) {
            snprintf(g_err, sizeof g_err, "import class %d (ages %d-%d) can be drawn but is empty", i, import_lo[i], import_hi[i]); free(e); return 1;
        }
    }
    e->row_len = RB_N_ATTRS * cfg->n_groups + RB_N_SCALARS;
    e->stats = (int32_t *)calloc((size_t)cfg->n_replicas * (cfg->max_days + 1) * e->row_len, sizeof(int32_t));
    e->sched = (rb_day_params *)calloc((size_t)cfg->max_days + 1, sizeof(rb_day_params));
    e->n_sched = cfg->max_days + 1;
    int bits = 1; while ((1u << bits) < (uint32_t)cfg->n_agents) bits++;
    e->feistel_half = (bits + 1) / 2;
    e->rep = (Replica *)calloc(cfg->n_replicas, sizeof(Replica));
    for (int ri = 0; ri < cfg->n_replicas; ri++) init_replica(e, ri, cfg->seed);
    *out = e;
    return 0;
}

static void init_replica(rb_engine *e, int ri, uint32_t seed) {
    const rb_config *cfg = &e->cfg;
    Replica *r = &e->rep[ri];
    if (r->agents) for (int32_t a = 0; a < cfg->n_agents; a++) free(r->agents[a].infectees);
    free(r->agents); free(r->order); free(r->perm); free(r->queue); free(r->newq);
    memset(r, 0, sizeof *r);
    r->seed = seed + (uint32_t)ri;
    r->agents = (Agent *)calloc(cfg->n_agents, sizeof(Agent));
    r->order = (int32_t *)malloc(sizeof(int32_t) * cfg->n_agents);
    r->perm = (int32_t *)malloc(sizeof(int32_t) * cfg->n_agents);
    philox(r->seed, KEY1, 0, 0, PU_PERM, 0, r->fkey);
    for (int age = 0; age < cfg->n_ages; age++) {
        r->counts[RB_A_SUSCEPTIBLE][age] = e->age_start[age + 1] - e->age_start[age];
        for (int32_t a = e->age_start[age]; a < e->age_start[age + 1]; a++) {
            Agent *p = &r->agents[a];
            p->age = (uint8_t)age; p->infector = -1; p->day_of_vaccination = -1; p->day_of_infection = -1;
            uint32_t s = feistel((uint32_t)a, (uint32_t)cfg->n_agents, e->feistel_half, r->fkey);
            r->perm[a] = (int32_t)s; r->order[s] = a;
        }
    }
    r->beds = r->avail_beds = cfg->hospital_beds;
    r->icu = r->avail_icu = cfg->icu_units;
    r->p_successful_tracing = 1.0f;
}

/* Population.set_initial_state, main.pyx:1452-1516 (called by Context.__init__ :1780-1781 while testing is still
 * NO_TESTING and day == 0).  ipc = {dead, in_icu, in_ward, confirmed_cases, incubating, ill, recovered}
 * (InitialPopulationCondition, calc/datasets.py:107-135).  Literal restatement, quirks included: people are drawn with
 * replacement and infected whatever their state (get_random_person :1518-1523), recovered_without_illness() equals
 * `incubating`, person_transfer_to_icu follows person_hospitalize unconditionally, all_detected is cleared for ages
 * 0..99 only and the confirmed cases are spread one per age. */
static void apply_initial_state(rb_engine *e, Replica *r) {
    const int32_t dead = e->ipc[0], in_icu = e->ipc[1], in_ward = e->ipc[2], confirmed = e->ipc[3];
    const int32_t incubating = e->ipc[4], ill = e->ipc[5], recovered = e->ipc[6];
    const int32_t were_ill = dead + recovered + in_icu + in_ward + ill, were_incubating = were_ill + incubating;
    const int32_t i_incubating = incubating, i_rws = i_incubating + (were_incubating - were_ill);
    const int32_t i_ill_at_home = i_rws + ill, i_dead = i_ill_at_home + dead, i_in_icu = i_dead + in_icu, i_in_ward = i_in_icu + in_ward;
    for (int32_t i = 0; i < were_incubating; i++) {
        uint32_t x[4];
        philox(r->seed, KEY1, (uint32_t)i, 0, PU_INIT, 0, x);
        const int32_t ai = (int32_t)(x[0] % (uint32_t)e->cfg.n_agents);
        Agent *p = &r->agents[ai];
        person_infect(e, r, ai, -1, 0, 0);
        if (i < i_incubating) continue;
        if (i < i_rws) { person_recover(r, p); continue; }
        person_become_ill(e, r, ai);
        if (i < i_ill_at_home) continue;
        if (i < i_dead) { person_die(r, p); continue; }
        if (i < i_in_icu) { person_hospitalize(e, r, ai); person_transfer_to_icu(e, r, ai); continue; }
        if (i < i_in_ward) { person_hospitalize(e, r, ai); continue; }
        person_recover(r, p);
    }
    for (int age = 0; age < 100 && age < e->cfg.n_ages; age++) r->counts[RB_A_ALL_DETECTED][age] = 0;
    for (int32_t i = 0; i < confirmed; i++) r->counts[RB_A_ALL_DETECTED][(100 + i) % 100] += 1;
}

int ro_set_initial_state(rb_engine *e, const int32_t *ipc7) {
    if (e->day != 0) { snprintf(g_err, sizeof g_err, "set_initial_state: only before the first step"); return 1; }
    memcpy(e->ipc, ipc7, sizeof e->ipc); e->has_ipc = 1;
    for (int ri = 0; ri < e->cfg.n_replicas; ri++) apply_initial_state(e, &e->rep[ri]);
    return 0;
}

int ro_reset(rb_engine *e, uint32_t seed) {
    e->cfg.seed = seed; e->day = 0;
    for (int ri = 0; ri < e->cfg.n_replicas; ri++) { init_replica(e, ri, seed); if (e->has_ipc) apply_initial_state(e, &e->rep[ri]); }
    return 0;
}

int ro_step(rb_engine *e, int32_t n_days);
int ro_step_profiled(rb_engine *e, int32_t n_days, float *ms) {
    for (int k = 0; k < RB_N_KERNELS; k++) ms[k] = 0;
    return ro_step(e, n_days);
}

void ro_destroy(rb_engine *e) {
    if (!e) return;
    for (int ri = 0; ri < e->cfg.n_replicas; ri++) {
        Replica *r = &e->rep[ri];
        for (int32_t a = 0; a < e->cfg.n_agents; a++) free(r->agents[a].infectees);
        free(r->agents); free(r->order); free(r->perm); free(r->queue); free(r->newq);
    }
    free(e->rep); free(e->tables); free(e->sched); free(e->stats); free(e);
}

int ro_set_contact_table(rb_engine *e, int32_t epoch, const int32_t *n_rows, const double *cum_p,
                         const int32_t *age_lo, const int32_t *age_hi, const uint8_t *place,
                         const float *mask_p, const double *nr_contacts, const double *ncontact_cdf) {
    if (epoch >= e->n_tables) {
        int n = epoch + 8;
        e->tables = (Table *)realloc(e->tables, sizeof(Table) * n);
        memset(e->tables + e->n_tables, 0, sizeof(Table) * (n - e->n_tables));
        e->n_tables = n;
    }
    /* a row that can be drawn needs somebody in its contact band (the person index is start + u32 % size; the reference
     * would divide by zero in get_person_from_age_range, main.pyx:1525-1535) */
    for (int age = 0; age < e->cfg.n_ages; age++) {
        if (n_rows[age] < 0 || n_rows[age] > RB_MAX_ROWS) { snprintf(g_err, sizeof g_err, "contact table: n_rows[%d] = %d", age, n_rows[age]); return 1; }
        if (e->age_start[age + 1] == e->age_start[age]) continue;
        for (int i = 0; i < n_rows[age]; i++) {
            int k = age * RB_MAX_ROWS + i;
            if (age_lo[k] < 0 || age_hi[k] >= e->cfg.n_ages || age_lo[k] > age_hi[k]) { snprintf(g_err, sizeof g_err, "contact table: bad band [%d, %d] (age %d row %d)", age_lo[k], age_hi[k], age, i); return 1; }
            double p = cum_p[k] - (i ? cum_p[k - 1] : 0.0);
            if ((p > 0.0 || i == n_rows[age] - 1) && e->age_start[age_hi[k] + 1] - e->age_start[age_lo[k]] <= 0) {
                snprintf(g_err, sizeof g_err, "contact table: row %d of age %d can be drawn (p = %g) but its contact band [%d, %d] is empty", i, age, p, age_lo[k], age_hi[k]);
                return 1;
            }
        }
    }
    Table *t = &e->tables[epoch];
    for (int age = 0; age < e->cfg.n_ages; age++) {
        t->n_rows[age] = n_rows[age];
        t->nr_contacts[age] = nr_contacts[age];
        memcpy(t->ncdf[age], ncontact_cdf + (size_t)age * 2 * RB_NCDF, sizeof(double) * 2 * RB_NCDF);
        for (int i = 0; i < n_rows[age]; i++) {
            int k = age * RB_MAX_ROWS + i;
            t->cum_p[age][i] = cum_p[k];
            t->start[age][i] = e->age_start[age_lo[k]];
            t->size[age][i] = e->age_start[age_hi[k] + 1] - e->age_start[age_lo[k]];
            t->place[age][i] = place[k];
            t->mask_p[age][i] = mask_p[k];
        }
    }
    t->set = 1;
    return 0;
}

int ro_set_schedule(rb_engine *e, int32_t day0, int32_t n, const rb_day_params *params) {
    if (day0 < 0 || day0 + n > e->n_sched) { snprintf(g_err, sizeof g_err, "schedule out of range"); return 1; }
    memcpy(e->sched + day0, params, sizeof(rb_day_params) * n);
    return 0;
}

int ro_snapshot(rb_engine *e) { for (int ri = 0; ri < e->cfg.n_replicas; ri++) snapshot(e, ri); return 0; }

int ro_step(rb_engine *e, int32_t n_days) {
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int d = 0; d < n_days; d++) {
        if (e->day >= e->cfg.max_days) { snprintf(g_err, sizeof g_err, "max_days exceeded"); return 1; }
        const rb_day_params *dp = &e->sched[e->day];
        if (dp->table_epoch >= e->n_tables || !e->tables[dp->table_epoch].set) { snprintf(g_err, sizeof g_err, "contact table %d not set", dp->table_epoch); return 1; }
        for (int ri = 0; ri < e->cfg.n_replicas; ri++) { snapshot(e, ri); iterate_replica(e, ri); }
        e->day += 1;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    e->last_ms = (float)((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
    return 0;
}
int ro_sync(rb_engine *e) { (void)e; return 0; }
int32_t ro_day(rb_engine *e) { return e->day; }
int ro_rng_block(int32_t words, uint32_t key, const uint32_t *ctr, uint32_t *out) {
    if (words == 4) { philox(key, KEY1, ctr[0], ctr[1], ctr[2], ctr[3], out); return 0; }
    snprintf(g_err, sizeof g_err, "ro_rng_block: words must be 4");
    return 1;
}
int32_t ro_row_len(rb_engine *e) { return e->row_len; }
int ro_read_stats(rb_engine *e, int32_t day0, int32_t n, int32_t *out) {
    for (int ri = 0; ri < e->cfg.n_replicas; ri++)
        memcpy(out + (size_t)ri * n * e->row_len,
               e->stats + ((size_t)ri * (e->cfg.max_days + 1) + day0) * e->row_len, sizeof(int32_t) * (size_t)n * e->row_len);
    return 0;
}
int ro_read_moments(rb_engine *e, int32_t day0, int32_t n, double *sum, double *sumsq) {
    for (int d = 0; d < n; d++)
        for (int col = 0; col < e->row_len; col++) {
            double s1 = 0.0, s2 = 0.0;
            for (int ri = 0; ri < e->cfg.n_replicas; ri++) {
                double v = (double)e->stats[((size_t)ri * (e->cfg.max_days + 1) + day0 + d) * e->row_len + col];
                s1 += v; s2 += v * v;
            }
            sum[(size_t)d * e->row_len + col] = s1; sumsq[(size_t)d * e->row_len + col] = s2;
        }
    return 0;
}
int ro_read_per_age(rb_engine *e, int32_t replica, int32_t attr, int32_t *out) {
    memcpy(out, e->rep[replica].counts[attr], sizeof(int32_t) * e->cfg.n_ages); return 0;
}
int ro_problem(rb_engine *e, int32_t *out) { for (int ri = 0; ri < e->cfg.n_replicas; ri++) out[ri] = e->rep[ri].problem; return 0; }
int ro_read_agents(rb_engine *e, int32_t replica, rb_agent *out) {
    Replica *r = &e->rep[replica];
    for (int32_t a = 0; a < e->cfg.n_agents; a++) {
        const Agent *p = &r->agents[a]; rb_agent *o = &out[a];
        o->infector = p->infector; o->n_infected = p->n_infected; o->days_left = p->days_left;
        o->day_of_illness = p->day_of_illness; o->day_of_vaccination = p->day_of_vaccination;
        o->state = p->state; o->severity = p->severity; o->variant = p->variant;
        o->flags = (uint8_t)(p->detected | (p->queued << 1) | (p->included << 2) | (p->has_list << 3));
        o->ward_days = p->ward_days; o->icu_days = p->icu_days;
    }
    return 0;
}
int ro_read_queue(rb_engine *e, int32_t replica, int32_t *out, int32_t cap, int32_t *n) {
    Replica *r = &e->rep[replica];
    *n = r->n_queue;
    for (int i = 0; i < r->n_queue && i < cap; i++) out[i] = r->queue[i];
    return 0;
}
int ro_read_available(rb_engine *e, int32_t replica, int32_t *o) { o[0] = e->rep[replica].avail_beds; o[1] = e->rep[replica].avail_icu; return 0; }
float ro_last_step_ms(rb_engine *e) { return e->last_ms; }
int64_t ro_launch_count(rb_engine *e) { (void)e; return 0; }

/* Context.sample, main.pyx:2047-2101: draws keyed (i, age, PU_SAMPLE | what << 8). */
int ro_sample(rb_engine *e, int32_t what, int32_t age, int32_t severity, int32_t n, int32_t *out) {
    const rb_variant *v = &e->variants[0];
    uint32_t seed = e->cfg.seed;
    int epoch = e->day > 0 ? e->sched[e->day - 1].table_epoch : e->sched[0].table_epoch;
    for (int32_t i = 0; i < n; i++) {
        uint32_t pu = PU_SAMPLE | ((uint32_t)what << 8);
        if (what == 0) {
            uint32_t x[4]; philox(seed, KEY1, (uint32_t)i, (uint32_t)age, pu, 0, x);
            double u = u01d(x[0], x[1]);
            const double *cdf = e->tables[epoch].ncdf[age][0];
            int lo = 0, hi = 100;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (u < cdf[mid]) hi = mid; else lo = mid + 1; }
            out[i] = lo;
        } else if (what == 1) {
            uint32_t x[4]; philox(seed, KEY1, (uint32_t)i, (uint32_t)age, pu, 0, x);
            out[i] = symptom_severity(v, age, u01f(x[0]), 0);
        } else if (what == 2) {
            out[i] = round_to_int(gamma_f(seed, (uint32_t)i, (uint32_t)age, pu, v->incubation_kappa, v->incubation_theta));
        } else {
            float T = severity == RB_FATAL ? gamma_f(seed, (uint32_t)i, (uint32_t)age, pu, v->onset_death_kappa, v->onset_death_theta)
                                           : gamma_f(seed, (uint32_t)i, (uint32_t)age, pu, v->onset_recovery_kappa, v->onset_recovery_theta);
            float f = 0.0f;
            if (what == 3) { f = T; if (severity != RB_ASYMPTOMATIC && severity != RB_MILD) f = f * v->ratio_before_hospitalisation; }
            else if (what == 4) { if (severity == RB_SEVERE) f = T * (1.0f - v->ratio_before_hospitalisation); else if (severity >= RB_CRITICAL) f = T * v->ratio_in_ward; }
            else if (what == 5) { if (severity >= RB_CRITICAL) f = ((1.0f - v->ratio_in_ward) - v->ratio_before_hospitalisation) * T; }
            else f = T;
            out[i] = round_to_int(f);
        }
    }
    return 0;
}
