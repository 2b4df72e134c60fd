// Counter-based random streams and deterministic float helpers (device side).
//
// Replaces cythonsim/simrandom.pyx:13-55 (one sequential PCG64) with Philox4x32-10 keyed on
// (seed; agent-or-ordinal, day, purpose|slot, iteration), so a fixed seed gives the same run whatever the
// kernel schedule.  Every function here is restated, operation for operation, in the CPU oracle
// (oracle/reina_oracle.c); the engine is compiled with -fmad=false so that both sides round identically.
#pragma once
#include <stdint.h>

enum { PU_START = 1, PU_NCONTACT, PU_CONTACT, PU_SEVERITY, PU_INCUB, PU_ONSET, PU_SEEK, PU_NOBED,
       PU_TRACE, PU_IMPORT, PU_PERM, PU_SAMPLE, PU_CONTACT2, PU_INIT };
#define RB_KEY1 0x5EEDB200u

struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ u32x4 philox(uint32_t k0, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k1 = RB_KEY1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    u32x4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// next_double (simrandom.pyx:24-26): 53 random bits in [0,1)
__host__ __device__ __forceinline__ double u01d(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
__host__ __device__ __forceinline__ float u01f(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
__host__ __device__ __forceinline__ float u01f_open(uint32_t x) { return (float)((x >> 8) + 1u) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float rb_logf(float x) {
    uint32_t b = __float_as_uint(x);
    int e = (int)(b >> 23) - 127;
    float m = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
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

__device__ __forceinline__ float rb_expf(float y) {
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
    return p * __uint_as_float((uint32_t)(ki + 127) << 23);
}

// Marsaglia polar method on words x,y of one Philox block; false when the pair is rejected.
__device__ __forceinline__ bool polar_normal(const u32x4 &x, float *z) {
    float u = 2.0f * u01f(x.x) - 1.0f, v = 2.0f * u01f(x.y) - 1.0f;
    float s = u * u + v * v;
    if (s >= 1.0f || s == 0.0f) return false;
    *z = u * sqrtf(-2.0f * rb_logf(s) / s);
    return true;
}

// random_gamma_f(kappa > 1, theta) -- Marsaglia-Tsang, one Philox block per trial (simrandom.pyx:46-55).
__device__ __forceinline__ float gamma_f(uint32_t seed, uint32_t c0, uint32_t c1, uint32_t purpose, float kappa, float theta) {
    float d = kappa - 0.333333333f;
    float c = 1.0f / sqrtf(9.0f * d);
    for (uint32_t it = 0;; it++) {
        u32x4 x = philox(seed, c0, c1, purpose, it);
        float z;
        if (!polar_normal(x, &z)) continue;
        float v = 1.0f + c * z;
        if (v <= 0.0f) continue;
        v = v * v * v;
        float u = u01f_open(x.z);
        float z2 = z * z;
        if (u < 1.0f - 0.0331f * (z2 * z2)) return (d * v) * theta;
        if (rb_logf(u) < 0.5f * z2 + d * ((1.0f - v) + rb_logf(v))) return (d * v) * theta;
    }
}

// random_lognormal(0, 0.5) (simrandom.pyx:41-44, used by get_nr_contacts main.pyx:1311)
__device__ __forceinline__ float lognormal_half(uint32_t seed, uint32_t c0, uint32_t c1, uint32_t purpose) {
    for (uint32_t it = 0;; it++) {
        u32x4 x = philox(seed, c0, c1, purpose, it);
        float z;
        if (polar_normal(x, &z)) return rb_expf(0.5f * z);
    }
}

// RandomPool.chance (simrandom.pyx:32-39)
__device__ __forceinline__ bool chance(double u, float p) {
    if (p == 1.0f) return true;
    if (p == 0.0f) return false;
    return u < (double)p;
}

__device__ __forceinline__ int round_to_int(float f) { return (int)(f + 0.5f); }   // main.pyx:773-774
__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Sweep order: keyed 4-round Feistel permutation of [0, n) with cycle walking (replaces the
// np.random.shuffle of agent indices, main.pyx:1436).
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
__host__ __device__ __forceinline__ uint32_t feistel(uint32_t a, uint32_t n, int half, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3) {
    uint32_t mask = (1u << half) - 1u, x = a;
    do {
        uint32_t L = x >> half, R = x & mask, t;
        t = L ^ (mix32(R ^ k0) & mask); L = R; R = t;
        t = L ^ (mix32(R ^ k1) & mask); L = R; R = t;
        t = L ^ (mix32(R ^ k2) & mask); L = R; R = t;
        t = L ^ (mix32(R ^ k3) & mask); L = R; R = t;
        x = (L << half) | R;
    } while (x >= n);
    return x;
}
