// vl_rng.cuh — counter-based Philox4x32-10 streams and exact samplers for the population update.
//
// The reference draws everything from an unseedable thread-local ChaCha12 (`rand::rng()`,
// src/simulation.rs:368,394,413,434,462,510), so only the *distributions* of its rand/rand_distr
// call sites can be matched:
//   Uniform<usize>(0,n)            src/simulation.rs:47,366      -> below()
//   rng.random::<f64>()            src/simulation.rs:64          -> u01()
//   Bernoulli(rho)                 src/simulation.rs:49          -> bernoulli() / geometric skipping
//   Binomial(L, mu)                src/simulation.rs:48          -> binomial()  (BINV | BTRS, exact)
//   WeightedIndex(row of 4)        src/simulation.rs:107-110     -> weighted4()
//   Poisson(lambda)                src/simulation.rs:148-151     -> poisson()   (Knuth | PTRS, exact)
// Every stream is keyed by (seed, compartment) and addressed by (index, generation, phase), so a
// draw does not depend on launch geometry, block scheduling or the number of GPUs.
#pragma once
#include <stdint.h>
#include <math.h>

namespace vl {

enum Phase : uint32_t {
    PH_INFECT = 1, PH_MUTCOUNT = 2, PH_MUTATE = 3, PH_RECOMB = 4, PH_OFFSPRING = 5, PH_PLAN = 6,
    PH_RESAMPLE = 7, PH_MIGRANT = 8, PH_TEST = 9, PH_PLAN_FILL = 10, PH_PLAN_TREE = 11, PH_RESAMPLE_RETRY = 12, PH_SHARD = 13
};

// ---- stateless block function: block `blk` of stream (index, generation, phase) under key (seed, compartment)
struct PhiloxKey { uint32_t k0, k1; };
__host__ __device__ inline PhiloxKey philox_key(uint64_t seed, uint32_t compartment) {
    uint64_t k = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(compartment + 1);
    k ^= k >> 31;
    return PhiloxKey{(uint32_t)k, (uint32_t)(k >> 32)};
}
__host__ __device__ inline void philox_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#ifdef __CUDA_ARCH__
    hi = __umulhi(a, b);
    lo = a * b;
#else
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}
__host__ __device__ inline uint4 philox_block(PhiloxKey key, uint64_t index, uint32_t generation, uint32_t phase, uint32_t blk) {
    uint32_t x0 = (uint32_t)index, x1 = (uint32_t)(index >> 32), x2 = generation, x3 = (phase << 24) + blk;
    uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, x0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, x2, hi1, lo1);
        uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(x0, x1, x2, x3);
}
// rand's random::<f64>() from two words: 53 random bits * 2^-53, in [0,1)
__host__ __device__ inline double u01_from(uint32_t lo, uint32_t hi) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
__host__ __device__ inline double u01_open_from(uint32_t lo, uint32_t hi) {
    return ((double)((((uint64_t)hi << 32) | lo) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// sequential generator over the same streams (general-purpose paths)
struct Philox {
    PhiloxKey key;
    uint64_t index;
    uint32_t generation, phase, blk;
    uint4 out;  // (selected by comparisons, never indexed: the generator stays in registers)
    int have;

    __host__ __device__ inline void block() {
        out = philox_block(key, index, generation, phase, blk);
        blk += 1;  // low 24 bits of the last counter word count blocks within a stream
        have = 4;
    }
    __host__ __device__ inline Philox(uint64_t seed, uint32_t compartment, uint32_t generation_, uint32_t phase_, uint64_t index_,
                                      uint32_t first_block = 0)
        : key(philox_key(seed, compartment)), index(index_), generation(generation_), phase(phase_), blk(first_block), have(0) {}
    __host__ __device__ inline uint32_t u32() {
        if (have == 0) block();
        const uint32_t r = have == 4 ? out.x : have == 3 ? out.y : have == 2 ? out.z : out.w;
        have -= 1;
        return r;
    }
    __host__ __device__ inline uint64_t u64() {
        uint64_t lo = u32();
        uint64_t hi = u32();
        return (hi << 32) | lo;
    }
    // rand's random::<f64>(): 53 random bits * 2^-53, in [0,1)
    __host__ __device__ inline double u01() { return (double)(u64() >> 11) * (1.0 / 9007199254740992.0); }
    // strictly inside (0,1): for samplers that take logs
    __host__ __device__ inline double u01_open() { return ((double)(u64() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    // Uniform integer in [0,n), n >= 1: widening multiply with rejection (unbiased)
    __host__ __device__ inline uint64_t below(uint64_t n) {
        if (n <= 0xFFFFFFFFull) {
            uint32_t n32 = (uint32_t)n;
            uint64_t m = (uint64_t)u32() * n32;
            uint32_t l = (uint32_t)m;
            if (l < n32) {
                uint32_t t = (uint32_t)(-n32) % n32;
                while (l < t) { m = (uint64_t)u32() * n32; l = (uint32_t)m; }
            }
            return m >> 32;
        }
        for (;;) {
            uint64_t x = u64();
#ifdef __CUDA_ARCH__
            uint64_t hi = __umul64hi(x, n);
#else
            uint64_t hi = (uint64_t)(((unsigned __int128)x * n) >> 64);
#endif
            uint64_t lo = x * n;
            if (lo >= n) return hi;
            uint64_t t = (0 - n) % n;
            if (lo >= t) return hi;
        }
    }
    __host__ __device__ inline bool bernoulli(double p) {
        if (p >= 1.0) return true;
        if (!(p > 0.0)) return false;
        return u64() < (uint64_t)(p * 18446744073709551616.0);
    }
};

// log(k!) - Stirling tail, as used by the transformed-rejection samplers (Hoermann 1993)
__host__ __device__ inline double stirling_tail(double k) {
    const double tail[10] = {0.0810614667953272, 0.0413406959554092, 0.0276779256849983, 0.02079067210376509,
                             0.0166446911898211, 0.0138761288230707, 0.0118967099458917, 0.0104112652619720,
                             0.00925546218271273, 0.00833056343336287};
    if (k <= 9.0) return tail[(int)k];
    double kp1sq = (k + 1.0) * (k + 1.0);
    return (1.0 / 12.0 - (1.0 / 360.0 - 1.0 / 1260.0 / kp1sq) / kp1sq) / (k + 1.0);
}

// Binomial(n, p), exact.  n*min(p,1-p) < 10: sequential inversion (the BINV branch rand_distr 0.5.1 also
// uses for small means); otherwise BTRS transformed rejection (same distribution as rand_distr's BTPE).
__host__ __device__ inline uint64_t binomial(Philox &g, uint64_t n, double p) {
    if (n == 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    const bool flipped = p > 0.5;
    if (flipped) p = 1.0 - p;
    const double dn = (double)n;
    uint64_t x;
    if (dn * p < 10.0) {
        const double q = 1.0 - p, s = p / q, a = (dn + 1.0) * s;
        const double r0 = exp(dn * log1p(-p));
        for (;;) {
            double r = r0, u = g.u01();
            uint64_t k = 0;
            while (u > r) {
                u -= r;
                k += 1;
                if (k > 200 || k > n) break;
                r *= a / (double)k - s;
            }
            if (k <= 200 && k <= n) { x = k; break; }
        }
    } else {
        const double spq = sqrt(dn * p * (1.0 - p));
        const double b = 1.15 + 2.53 * spq, a = -0.0873 + 0.0248 * b + 0.01 * p, c = dn * p + 0.5;
        const double vr = 0.92 - 4.2 / b, r = p / (1.0 - p), alpha = (2.83 + 5.1 / b) * spq;
        const double m = floor((dn + 1.0) * p);
        for (;;) {
            double u = g.u01_open() - 0.5, v = g.u01_open();
            double us = 0.5 - fabs(u);
            double k = floor((2.0 * a / us + b) * u + c);
            if (us >= 0.07 && v <= vr) { x = (uint64_t)k; break; }
            if (k < 0.0 || k > dn) continue;
            v = log(v * alpha / (a / (us * us) + b));
            double upper = (m + 0.5) * log((m + 1.0) / (r * (dn - m + 1.0))) +
                           (dn + 1.0) * log((dn - m + 1.0) / (dn - k + 1.0)) +
                           (k + 0.5) * log(r * (dn - k + 1.0) / (k + 1.0)) + stirling_tail(m) +
                           stirling_tail(dn - m) - stirling_tail(k) - stirling_tail(dn - k);
            if (v <= upper) { x = (uint64_t)k; break; }
        }
    }
    return flipped ? n - x : x;
}

// Poisson(lambda) for finite lambda > 0, exact.  lambda < 12: Knuth's product method (the branch
// rand_distr 0.5.1 uses there); otherwise PTRS transformed rejection (same distribution as its
// Cauchy-envelope rejection).
__host__ __device__ inline uint64_t poisson(Philox &g, double lambda) {
    if (lambda < 12.0) {
        const double limit = exp(-lambda);
        uint64_t k = 0;
        double p = g.u01();
        while (p > limit) { p *= g.u01(); k += 1; }
        return k;
    }
    const double slam = sqrt(lambda), loglam = log(lambda);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        double u = g.u01_open() - 0.5, v = g.u01_open();
        double us = 0.5 - fabs(u);
        double k = floor((2.0 * a / us + b) * u + lambda + 0.43);
        if (us >= 0.07 && v <= vr) return (uint64_t)k;
        if (k < 0.0 || (us < 0.013 && v > us)) continue;
        if (log(v) + log(invalpha) - log(a / (us * us) + b) <= -lambda + k * loglam - lgamma(k + 1.0)) return (uint64_t)k;
    }
}

// rand WeightedIndex<f64> over one substitution row: cumulative weights, chosen ~ U[0,total),
// index = partition_point(cum <= chosen)
__host__ __device__ inline int weighted4(Philox &g, const double *w) {
    double c0 = w[0], c1 = c0 + w[1], c2 = c1 + w[2], total = c2 + w[3];
    double chosen = g.u01() * total;
    int idx = 0;
    if (c0 <= chosen) idx = 1;
    if (c1 <= chosen) idx = 2;
    if (c2 <= chosen) idx = 3;
    return idx;
}


// ---- hot-path samplers: one Philox block per decision wherever possible -------------------------------------
// Poisson(lambda >= 12): PTRS transformed rejection (Hoermann 1993), one block per attempt (same law as
// rand_distr 0.5.1's rejection branch).
__device__ __noinline__ uint64_t poisson_ptrs(PhiloxKey key, uint64_t index, uint32_t generation, uint32_t phase, double lambda) {
    const double slam = sqrt(lambda);
    const double bb = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * bb;
    const double vr = 0.9277 - 3.6224 / (bb - 2.0);
    for (uint32_t attempt = 0;; attempt++) {
        const uint4 b = philox_block(key, index, generation, phase, attempt);
        const double u = u01_open_from(b.x, b.y) - 0.5, v = u01_open_from(b.z, b.w);
        const double us = 0.5 - fabs(u);
        const double k = floor((2.0 * a / us + bb) * u + lambda + 0.43);
        if (us >= 0.07 && v <= vr) return (uint64_t)k;
        if (k < 0.0 || (us < 0.013 && v > us)) continue;
        const double invalpha = 1.1239 + 1.1328 / (bb - 3.4);
        if (log(v) + log(invalpha) - log(a / (us * us) + bb) <= -lambda + k * log(lambda) - lgamma(k + 1.0)) return (uint64_t)k;
    }
}
// Poisson(lambda), exact.  lambda < 12: inversion of ONE uniform (words x,y of block 0) by sequential search —
// same law as rand_distr 0.5.1's Knuth product branch at a fraction of the random words.  `rcp` is a table of
// 1/k for k < POISSON_RCP (the term ratio lambda/k as a multiplication; distribution unchanged to 1 ulp of a term).
constexpr int POISSON_RCP = 64;
// 1/k, correctly rounded (what 1.0 / (double)k and __frcp_rn((float)k) return), so no kernel spends divisions on them
__constant__ double c_rcp64[POISSON_RCP] = {0.0, 1.0, 0.5, 0.3333333333333333, 0.25, 0.2, 0.16666666666666666, 0.14285714285714285, 0.125, 0.1111111111111111, 0.1, 0.09090909090909091, 0.08333333333333333, 0.07692307692307693, 0.07142857142857142, 0.06666666666666667, 0.0625, 0.058823529411764705, 0.05555555555555555, 0.05263157894736842, 0.05, 0.047619047619047616, 0.045454545454545456, 0.043478260869565216, 0.041666666666666664, 0.04, 0.038461538461538464, 0.037037037037037035, 0.03571428571428571, 0.034482758620689655, 0.03333333333333333, 0.03225806451612903, 0.03125, 0.030303030303030304, 0.029411764705882353, 0.02857142857142857, 0.027777777777777776, 0.02702702702702703, 0.02631578947368421, 0.02564102564102564, 0.025, 0.024390243902439025, 0.023809523809523808, 0.023255813953488372, 0.022727272727272728, 0.022222222222222223, 0.021739130434782608, 0.02127659574468085, 0.020833333333333332, 0.02040816326530612, 0.02, 0.0196078431372549, 0.019230769230769232, 0.018867924528301886, 0.018518518518518517, 0.01818181818181818, 0.017857142857142856, 0.017543859649122806, 0.017241379310344827, 0.01694915254237288, 0.016666666666666666, 0.01639344262295082, 0.016129032258064516, 0.015873015873015872};
__constant__ float c_rcp32[48] = {0.0f, 1.0f, 0.5f, 0.3333333432674408f, 0.25f, 0.20000000298023224f, 0.1666666716337204f, 0.1428571492433548f, 0.125f, 0.1111111119389534f, 0.10000000149011612f, 0.09090909361839294f, 0.0833333358168602f, 0.07692307978868484f, 0.0714285746216774f, 0.06666667014360428f, 0.0625f, 0.05882352963089943f, 0.0555555559694767f, 0.05263157933950424f, 0.05000000074505806f, 0.0476190485060215f, 0.04545454680919647f, 0.043478261679410934f, 0.0416666679084301f, 0.03999999910593033f, 0.03846153989434242f, 0.03703703731298447f, 0.0357142873108387f, 0.03448275849223137f, 0.03333333507180214f, 0.032258063554763794f, 0.03125f, 0.03030303120613098f, 0.029411764815449715f, 0.02857142873108387f, 0.02777777798473835f, 0.027027027681469917f, 0.02631578966975212f, 0.025641025975346565f, 0.02500000037252903f, 0.024390242993831635f, 0.02380952425301075f, 0.023255813866853714f, 0.022727273404598236f, 0.02222222276031971f, 0.021739130839705467f, 0.021276595070958138f};
template <bool SMALL_ONLY = false>
__device__ __forceinline__ uint64_t poisson_draw(PhiloxKey key, uint64_t index, uint32_t generation, uint32_t phase, double lambda,
                                                 const double *rcp) {
    if (SMALL_ONLY || lambda < 12.0) {
        const uint4 b = philox_block(key, index, generation, phase, 0);
        const double u = u01_from(b.x, b.y);
        double p = exp(-lambda), cdf = p;
        uint32_t k = 0;
        while (u >= cdf && k < 256) {
            k += 1;
            p = (k < POISSON_RCP) ? p * (lambda * rcp[k]) : p * lambda / (double)k;
            cdf += p;
        }
        return k;
    }
    if (SMALL_ONLY) return 0;
    return poisson_ptrs(key, index, generation, phase, lambda);
}
// The same inversion, decided in f32 whenever that is provably the f64 decision: the search runs on f32 copies of u
// and of the CDF.  For lambda < 12 and k <= 46 the f32 CDF is within 5e-6 of the f64 one: __expf(-lambda) is good to
// 1.2e-6 relative (7e-7 from the rounded argument, 5e-7 from ex2.approx), every term multiplies in three roundings
// (1.8e-7), so a term is good to 4e-6 relative after 16 steps and the terms sum to at most 1; the additions contribute
// 46 x 6e-8.  If u stays more than POISSON_GUARD away from every CDF value it is compared with, both searches stop at
// the same k.  Otherwise (about 5 draws in 10^4) the f64 search above decides.  The result is the
// f64 inversion's for every (u, lambda): vl_test_poisson_fast checks that draw for draw.
constexpr float POISSON_GUARD = 3.0e-5f;  // 6x the 5e-6 bound derived above
__device__ __forceinline__ uint32_t poisson_inversion_f64(double u, double lambda, const double *rcp) {
    double p = exp(-lambda), cdf = p;
    uint32_t k = 0;
    while (u >= cdf && k < 256) {
        k += 1;
        p = (k < POISSON_RCP) ? p * (lambda * rcp[k]) : p * lambda / (double)k;
        cdf += p;
    }
    return k;
}
// `kmax` is the same for every lane (it comes from the largest mean a launch can see): all lanes walk the CDF up to
// kmax without a data-dependent branch — in a warp the search would run to the slowest lane's answer anyway — and the
// answer is the number of CDF values that u has reached.
__device__ __forceinline__ uint32_t poisson_draw_small_fast(PhiloxKey key, uint64_t index, uint32_t generation, uint32_t phase, double lambda,
                                                            const double *rcp, uint32_t kmax) {
    const uint4 b = philox_block(key, index, generation, phase, 0);
    const double u = u01_from(b.x, b.y);
    const float uf = (float)u, lf = (float)lambda;
    float p = __expf(-lf), cdf = p, mind = fabsf(uf - cdf);
    uint32_t k = uf >= cdf ? 1u : 0u;
    for (uint32_t j = 1; j <= kmax; j++) {
        p *= lf * c_rcp32[j];  // j is uniform: one constant-bank operand for the warp
        cdf += p;
        mind = fminf(mind, fabsf(uf - cdf));
        k += uf >= cdf ? 1u : 0u;
    }
    if (k <= kmax && mind > POISSON_GUARD) return k;  // u < CDF(kmax) and clear of every CDF value it was compared with
    return poisson_inversion_f64(u, lambda, rcp);
}
// kmax for means up to `max_mean` (< 12): beyond it lies less than 1e-3 of the mass, those draws take the f64 search
__host__ __device__ inline uint32_t poisson_kmax(double max_mean) {
    double k = max_mean + 3.2 * sqrt(max_mean > 0.0 ? max_mean : 0.0) + 2.0;
    return k < 46.0 ? (uint32_t)k : 46u;  // < the 48 entries of c_rcp32
}
// fills a CTA's shared reciprocal table (call by all threads, then __syncthreads())
__device__ __forceinline__ void poisson_rcp_init(double *rcp) {
    for (int k = threadIdx.x; k < POISSON_RCP; k += blockDim.x) rcp[k] = c_rcp64[k];
}

// Binomial(n, p) in the small-mean regime n*p < 10 with host-precomputed constants q^n, s = p/q, a = (n+1)s:
// inversion of one uniform (the BINV branch of rand_distr 0.5.1).  `u` is consumed; a restart (k beyond the
// search bound, probability < 1e-100) draws from the stream's later blocks.
constexpr int BINV_STEPS = 16;
constexpr int BINV_THRESHOLDS = 16;
struct BinvConst {
    double qn, s, a; uint64_t n; int ok; double step[BINV_STEPS]; /* a/k - s for k < BINV_STEPS */
    // thr[k] = the smallest 53-bit integer U for which the inversion of u = U * 2^-53 returns more than k: the count
    // is the number of thresholds U reaches.  Found on the host by bisection over the very search the device runs
    // (binv_count below, IEEE operations only), so comparing integers gives the inversion's answer bit for bit.
    uint64_t thr[BINV_THRESHOLDS];
};
// first pass of the inversion (no restart): steps taken for this uniform
__host__ __device__ inline uint32_t binv_count(const BinvConst &c, double u) {
    double r = c.qn;
    uint32_t k = 0;
    while (u > r) {
        u -= r;
        k += 1;
        if (k > 200 || k > c.n) break;
        r *= (k < (uint32_t)BINV_STEPS) ? c.step[k] : c.a / (double)k - c.s;
    }
    return k;
}
__host__ inline BinvConst make_binv(uint64_t n, double p) {
    BinvConst c;
    c.n = n;
    c.ok = (p > 0.0 && p <= 0.5 && (double)n * p < 10.0) ? 1 : 0;
    const double q = 1.0 - p;
    c.qn = c.ok ? exp((double)n * log1p(-p)) : 0.0;
    c.s = p / q;
    c.a = ((double)n + 1.0) * c.s;
    for (int k = 0; k < BINV_STEPS; k++) c.step[k] = k ? c.a / (double)k - c.s : 0.0;
    for (int k = 0; k < BINV_THRESHOLDS; k++) {
        // binv_count is non-decreasing in u (every step subtracts the same constants): bisect for the first U with count > k
        uint64_t lo = 0, hi = 1ull << 53;  // hi = "never"
        while (c.ok && lo < hi) {
            const uint64_t mid = lo + ((hi - lo) >> 1);
            if (binv_count(c, (double)mid * (1.0 / 9007199254740992.0)) > (uint32_t)k) hi = mid; else lo = mid + 1;
        }
        c.thr[k] = c.ok ? lo : (1ull << 53);
    }
    return c;
}
__device__ __forceinline__ uint32_t binv_draw(const BinvConst &c, double u, PhiloxKey key, uint64_t index, uint32_t generation, uint32_t phase) {
    uint32_t blk = 1;
    for (;;) {
        double r = c.qn;
        uint32_t k = 0;
        while (u > r) {
            u -= r;
            k += 1;
            if (k > 200 || k > c.n) break;
            r *= (k < (uint32_t)BINV_STEPS) ? c.step[k] : c.a / (double)k - c.s;
        }
        if (k <= 200 && k <= c.n) return k;
        const uint4 b = philox_block(key, index, generation, phase, blk++);
        u = u01_from(b.x, b.y);
    }
}

}  // namespace vl
