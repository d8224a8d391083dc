// vl_state.cuh — device-resident state of one compartment and the haplotype record format.
//
// Layout in HBM (structure of arrays, one set per vl_sim):
//   population     hap_id[n]            u32   id of each infectant's haplotype      (Population.population,
//                                                                                     src/core/population.rs:156-163)
//   host map       host_id[n]           u32   Option<host> with NONE sentinel        (HostMapBuffer.infections,
//                  offsets[H+1]         u32   CSR by host                             src/core/hosts.rs:133-192)
//                  rec[n_infected]      uint2 (infectant index, haplotype id) in CSR order: HostMapBuffer.hosts with
//                                             the haplotype id carried along, so everything downstream of the
//                                             counting sort streams over host-contiguous records
//   replication    lam[p] f64 (Poisson mean of the record at CSR position p), offspring[p] u32, tile_sum[] u64
//   haplotypes     head[id] = {base list offset u64, length u32, delta count | merged length, raw0, fit0, 8 delta entries}
//                  (64 B) + h_raw[provider][id], h_fit[provider][id] f64 + arena of u32 entries
//
// A haplotype record is the *flattened* form of the reference's delta chain (Wildtype / Mutant{changes} /
// Recombinant{left,right,window}, src/core/haplotype.rs:45-56,241-288): the position-sorted list of sites where
// the haplotype differs from the wildtype, one u32 entry per site:
//       entry = pos << 8 | g << 4 | m
//   g = what Haplotype::get_base(pos) returns            (first matching change wins,  haplotype.rs:661-667)
//   m = what Haplotype::get_mutations()[pos] holds, 4 if absent (last change wins, back-mutation to the
//       wildtype base removes the key,                                                   haplotype.rs:686-705)
// The two views differ only after one mutation event hit the same site twice (sites are drawn with
// replacement, src/simulation.rs:94-103); keeping both makes the flattened record answer every query the
// chain answers.  Raw fitness per provider is computed once when the record is created, exactly as
// FitnessProvider::get_fitness would on first touch (src/providers/fitness.rs:212-228), and memoised in
// the haplotype's metadata (AttributeSet, src/core/attributes.rs:196-217).
#pragma once
#include <stdint.h>
#include "../../include/virolution_gpu.h"
#include "vl_rng.cuh"

namespace vl {

constexpr uint32_t NONE32 = 0xFFFFFFFFu;
constexpr int MAX_SPECS = 8;
constexpr int MAX_PROVIDERS = 16;
constexpr uint32_t RESAMPLE_TILE = 1024;  // infectants per resample tile (one CTA stages one tile in shared memory)
constexpr uint32_t HOST_TILE_CAP = 2048;  // records a host tile of the fused replicate kernel may hold
constexpr int MAX_PARTS = 16;              // subsamples of one offspring map that one plan + one resample pass can draw together
constexpr uint32_t HT_MAX_HOSTS = HOST_TILE_CAP;   // most hosts a host-range tile may span (DevState::host_tiles holds the actual number)
constexpr uint32_t HT_MIN_HOSTS = 256;    // below this the fused replicate kernel is not worth launching
constexpr double HT_SPREAD_CTAS = 148.0;  // tiles of a small host population are shrunk until there are about two per SM of a B200
// error bits raised by kernels (sticky until vl_* reports them)
enum DevError : uint32_t {
    DE_HAP_CAPACITY = 1u,       // haplotype table full
    DE_ARENA_CAPACITY = 2u,     // arena full
    DE_ALL_ZERO_WEIGHTS = 4u,   // WeightedAliasIndex::new panics (src/core/population.rs:386)
    DE_UNDEFINED_OFFSPRING = 8u,// Poisson::new failed on a non-zero fitness (src/simulation.rs:148-151)
    DE_SAME_BASE = 16u,         // assert_ne!(from, to) (src/core/haplotype.rs:250)
    DE_INF_CAPACITY = 32u,      // population buffer too small
    DE_OFFSPRING_RANGE = 64u,   // offspring count does not fit u32
    DE_REPLAY = 128u,           // malformed replay log
    DE_BUCKET_OVERFLOW = 256u,  // a host bucket received 12 standard deviations more records than its mean
    DE_TILE_OVERFLOW = 512u,    // a host tile holds more records than a CTA stages (uniform infection: > 8 sd)
    DE_EXCHANGE_TIMEOUT = 1024u, // a peer's migrants did not arrive in time (vl_exchange_step)
    DE_PLAN_TIMEOUT = 2048u     // the CTAs of the resample plan did not all become resident (grid barrier)
};
constexpr uint32_t BUCKET_SHIFT = 13;
constexpr uint32_t BUCKET_HOSTS = 1u << BUCKET_SHIFT;  // hosts per bucket: their histogram lives in shared memory
// records reserved per bucket: mean + 12 sd of Binomial(n, BUCKET_HOSTS / H) + slack (never reached in practice)
__host__ __device__ inline uint64_t bucket_stride(uint64_t n, uint64_t H) {
    const double m = (double)n * (double)BUCKET_HOSTS / (double)(H ? H : 1);
    return (uint64_t)(m + 12.0 * sqrt(m) + 64.0);
}

struct EpiDir {  // one direction of a symmetric pair entry (src/core/fitness/epistasis.rs:44-58)
    uint32_t k1, k2;  // (pos << 2) | base
    double v;
};

struct DevProvider {
    int32_t kind;  // vl_fitness_kind
    int32_t util_kind;
    double upper, hill_k, hill_n;
    double alg_factor, alg_uf;  // Algebraic: factor = 1/(upper-1) and upper*factor, rounded exactly as the reference forms them
    const double *table;  // L*4
    const EpiDir *epi;    // sorted by (k1, k2), duplicates resolved (last insert wins)
    const uint32_t *epi_start;  // L*4 + 1: first entry of every key's row (row of key k = [epi_start[k], epi_start[k + 1])), or nullptr
    uint32_t n_epi;
    uint32_t n_keys;      // L*4 when epi_start is set
};

struct DevSpec {
    uint64_t lo, hi;
    int32_t repro, infect;  // provider index or -1
};

// One haplotype record = one 64-byte line.  The reference stores a delta chain (Mutant{parent, changes},
// src/core/haplotype.rs:45-56,570-594): a mutation costs O(1) whatever the depth, a base lookup walks the chain
// (:661-667).  Here a record is a BASE LIST (position-sorted flattened entries in the arena, shared with the ancestors it
// was inherited from) plus up to HEAD_DELTA DELTA ENTRIES inline (position-sorted, each overriding the base entry at its
// position; m == ENTRY_TOMBSTONE removes it).  A mutation copies the 64-byte line and merges its changes into the
// delta — O(1) in the haplotype's length; when the delta would overflow, the record is flattened into a fresh base list
// (one O(length) copy per HEAD_DELTA mutations of a lineage, the stand-in for Mutant::try_merge_node, :743-784).
// Lookups scan the delta, then search the base list by interpolation (sites are drawn uniformly, so the list is
// uniformly spaced: ~2 probes).  The compaction (k_gc_*) flattens every live record.
constexpr uint32_t HEAD_DELTA = 8;
constexpr uint32_t ENTRY_TOMBSTONE = 5;  // m field of a delta entry that removes the base entry at its position
struct alignas(64) HapHead {
    uint64_t off;       // first arena word of the base list
    uint32_t len;       // base entries
    uint32_t nd_mlen;   // delta entries (low 8 bits) | entries of the merged record << 8
    double raw0;        // raw fitness under provider 0 (copy of h_raw[0][id]: the mutate kernel reads ONE line per parent)
    double fit0;        // mapped fitness under provider 0 (copy of h_fit[0][id])
    uint32_t d[HEAD_DELTA];
};
static_assert(sizeof(HapHead) == 64, "one line per haplotype");
__host__ __device__ inline uint32_t head_nd(const HapHead &h) { return h.nd_mlen & 0xFFu; }
__host__ __device__ inline uint32_t head_mlen(const HapHead &h) { return h.nd_mlen >> 8; }

struct DevState {
    // ---- configuration (BasicHost.parameters are the values captured at vl_create, the simulation's
    //      are updatable through vl_set_parameters; src/simulation.rs:41-52 vs :351-363)
    uint32_t L;
    uint32_t n_specs, n_providers;
    uint32_t compartment;
    uint64_t seed;
    double host_mutation_rate, host_recombination_rate, host_r0;
    double host_subst[16];
    uint64_t H;            // simulation.parameters.host_population_size
    uint64_t n_hits;
    double dilution;
    uint64_t max_population;
    DevSpec specs[MAX_SPECS];
    DevProvider prov[MAX_PROVIDERS];
    const uint8_t *wildtype;

    // ---- population
    uint64_t n;            // current number of infectants
    uint64_t n_next;       // size of the population being assembled
    uint64_t cap_inf;
    uint32_t *hap_id;      // current
    uint32_t *hap_id_next; // being assembled
    uint32_t generation;
    uint32_t error;
    uint64_t infectant_generations;

    // ---- host map (CSR positions p in [0, n_sorted))
    uint32_t *host_id, *rank, *offsets;
    uint2 *rec;            // (infectant index, haplotype id) per CSR position
    uint64_t n_sorted;     // number of records = infected infectants (offsets[H])
    uint32_t *heavy_hosts; // worklist of hosts whose slice exceeds the thread tier
    uint32_t *medium_hosts; // worklist of hosts with 5 .. THREAD_TIER members
    uint32_t n_medium;     // n_medium, n_heavy, fused_failed are cleared together before a replicate phase
    uint32_t n_heavy;
    uint32_t fused_failed; // a tile of the fused replicate kernel did not fit: the general kernels redo the phase
    uint32_t host_tiles;   // > 0: resample tiles are ranges of that many hosts, 0: ranges of RESAMPLE_TILE CSR positions
    uint32_t ht_planned;   // host_tile_size(n, H) of the current population, set before the fused replicate kernel runs
    uint32_t fast_infect;  // specs cover [0,H), no infective provider, n_hits == 1: every draw infects, hosts uniform
    // ---- bucketed counting sort (hosts split into buckets of BUCKET_HOSTS consecutive hosts)
    uint32_t n_buckets;
    uint32_t _pad2;
    uint32_t *bucket_fill; // records appended per bucket
    uint32_t *bucket_pos;  // n_buckets + 1: first CSR position of every bucket
    uint4 *tmp_rec;        // bucket-partitioned records (index, haplotype, host, -), bucket b at [b * stride, b * stride + fill)
    uint64_t cap_tmp;

    // ---- device-resident generation (vl_fused.cuh): per-host fitness sums by f64 atomics, no host map
    double *hsum[2];       // H sums each: generation g accumulates into hsum[hsum_parity] and clears the other one
    uint32_t hsum_parity;
    uint32_t _pad5;
    double *pop_fit, *pop_fit_next;  // mapped reproductive fitness per infectant, parallel to hap_id / hap_id_next (valid when the host says so)

    // ---- mutation worklist
    uint4 *mut_list;       // (infectant index, n_mut, position in the bucket-partitioned records, -) where n_mut > 0
    uint32_t *big_list;    // worklist positions of the events with more sites than the shared scratch holds (k_mutate, deferred pass)
    uint32_t n_mut_list;
    uint32_t mut_ctas_done;  // CTAs of the running k_mutate that have left (the last one publishes n_hap)
    uint32_t mut_own;        // worklist items [0, mut_own) are mutated ahead of their generation (vl_exchange_run: while the migrants travel)
    uint32_t tail_done;      // CTAs of the running k_resample_pos that have left (the last one installs the new population)
    uint32_t n_big;          // entries of big_list (zero between generations)
    uint32_t big_pad_;
    unsigned long long mut_own_base;  // ... with record ids mut_own_base + worklist position (reserved by k_ex_own_items)
    uint32_t _pad0;
    BinvConst binv;        // Binomial(L, mu) inversion constants (captured BasicHost parameters)

    // ---- replication
    double *lam;           // per CSR position: mapped fitness f, later the Poisson mean
    double *hs;            // per CSR position: fitness sum of the record's host
    uint32_t *offspring;   // per CSR position
    uint64_t *tile_sum;    // per resample tile
    uint64_t *tile_tree;   // segment tree over tile sums (fallback plan) / inclusive prefix of tile sums
    uint64_t *tile_cnt;    // draws per tile, one row of tile_stride entries per part
    uint64_t *tile_node;   // draws per tree node (fallback plan)
    uint64_t *tile_base;   // exclusive scan of draws per tile, per part
    uint64_t offspring_total;
    double target_size;    // of the last plan
    uint32_t plan_sync;    // grid barrier counter of the plan kernel; plan_sync .. plan_acc are zero between launches: the last
    uint32_t plan_exit;    // CTA to leave the kernel (plan_exit counts them) clears them for the next launch
    uint64_t plan_acc[4][MAX_PARTS];   // draws made per Poissonisation stage and part, summed over the plan CTAs
    uint64_t plan_part[32][MAX_PARTS]; // draws per CTA slice of tiles and part
    uint64_t plan_size[MAX_PARTS];     // sample size of every part of the last plan
    uint64_t plan_amount[MAX_PARTS];   // sample sizes decided on the device (sharded compartment), PlanArgs::use_amount == 2
    uint64_t tile_stride;              // tile_cnt / tile_base hold MAX_PARTS rows of this many tiles

    // ---- haplotype table (double-buffered for compaction; the live set is swapped on device)
    uint64_t cap_hap, cap_arena;
    unsigned long long n_hap;
    unsigned long long arena_top;
    uint32_t gc_active;
    uint32_t _pad1;
    HapHead *head, *head_alt;  // 64 bytes per haplotype: base list, inline delta, fitness under provider 0
    double *h_raw[MAX_PROVIDERS], *h_raw_alt[MAX_PROVIDERS];  // memoised raw fitness per provider
    double *h_fit[MAX_PROVIDERS], *h_fit_alt[MAX_PROVIDERS];  // utility(raw): what readers of get_or_compute see (attributes.rs:220-241)
    uint32_t *arena, *arena_alt;
    // compaction / dedup scratch, one entry per 32 records: live bit map, live records and live (merged) entries per word,
    // exclusive scans of the two
    uint32_t *gc_bits, *gc_wcnt, *gc_wlen, *gc_cbase;
    uint64_t *gc_lbase;
    uint64_t gc_nw;        // words of the bit map in use: ceil(n_hap / 32), written by k_gc_clear for the scans
    uint64_t gc_hap_threshold, gc_arena_threshold;
    uint64_t gc_runs;
    uint64_t kernels_launched;  // maintained by the host, mirrored here for export
    uint64_t dbg[8];            // %globaltimer stamps of the plan kernel's phases (profiling aid)
};

// hosts per tile of the fused replicate kernel: expected records per tile 8 standard deviations below what a CTA
// stages (RESAMPLE_TILE) when n infectants are spread uniformly over H hosts; 0 = host tiles cannot be used.
// The kernel evaluates it on the device from the population size, so the tiling (and with it the resampling streams)
// is a function of the simulation state alone; the host evaluates it on an upper bound of n to size grids.
__host__ __device__ inline uint32_t host_tile_size(uint64_t n, uint64_t H) {
    if (H == 0) return 0;
    const double mean_max = 1700.0;  // 1700 + 8 * sqrt(1700) < 2048
    double ht = floor(mean_max * (double)H / (double)(n ? n : 1));
    if (ht > (double)HT_MAX_HOSTS) ht = (double)HT_MAX_HOSTS;
    if (ht > (double)H) ht = (double)H;
    if (ht < (double)HT_MIN_HOSTS) return 0u;
    // small host populations: rather many small tiles than a handful of full ones (a tile is one CTA; a generation of
    // 10^4 infectants should still spread over the SMs)
    const double spread = ceil((double)H / (2.0 * HT_SPREAD_CTAS));
    if (ht > spread) ht = spread > (double)HT_MIN_HOSTS ? spread : (double)HT_MIN_HOSTS;
    // a power of two (HT_MIN_HOSTS and HT_MAX_HOSTS are): the tile of a host is a shift, and every path that tiles by
    // hosts — the trait-level fused replicate kernel and the tile-strided generation — cuts the same tiles
    uint32_t p = HT_MIN_HOSTS;
    while ((double)(2u * p) <= ht) p *= 2u;
    return p;
}

// ---- resample tiles: a tile is what one CTA stages in shared memory, either RESAMPLE_TILE consecutive CSR positions
// or the records of DevState::host_tiles consecutive hosts (what the fused replicate kernel produces: whole hosts per CTA)
__device__ __forceinline__ uint64_t tile_count(const DevState *st) {
    const uint64_t ht = st->host_tiles;
    return ht ? (st->H + ht - 1) / ht : (st->n_sorted + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
}
__device__ __forceinline__ void tile_range(const DevState *st, uint64_t tile, uint64_t &start, uint32_t &len) {
    if (st->host_tiles) {
        const uint64_t ht = st->host_tiles;
        const uint64_t h0 = tile * ht, h1 = h0 + ht < st->H ? h0 + ht : st->H;
        start = st->offsets[h0];
        len = st->offsets[h1] - (uint32_t)start;
    } else {
        start = tile * RESAMPLE_TILE;
        const uint64_t rest = st->n_sorted - start;
        len = rest < RESAMPLE_TILE ? (uint32_t)rest : RESAMPLE_TILE;
    }
}

// memoised raw fitness of haplotype `id` under provider p (AttributeSet::get_or_compute_raw, src/core/attributes.rs:196-217)
// (kept apart from the 64-byte heads: the per-infectant fitness gather wants the densest possible array)
__host__ __device__ inline double &hap_raw(DevState *st, uint32_t p, uint64_t id) { return st->h_raw[p][id]; }
__host__ __device__ inline const double &hap_raw(const DevState *st, uint32_t p, uint64_t id) { return st->h_raw[p][id]; }
__host__ __device__ inline double &hap_raw_alt(DevState *st, uint32_t p, uint64_t id) { return st->h_raw_alt[p][id]; }

// ---- entry helpers ---------------------------------------------------------------------------------
__host__ __device__ inline uint32_t make_entry(uint32_t pos, uint32_t g, uint32_t m) { return (pos << 8) | (g << 4) | m; }
__host__ __device__ inline uint32_t entry_pos(uint32_t e) { return e >> 8; }
__host__ __device__ inline uint32_t entry_g(uint32_t e) { return (e >> 4) & 0xF; }
__host__ __device__ inline uint32_t entry_m(uint32_t e) { return e & 0xF; }  // 4 = not in the mutation set

__host__ __device__ inline bool entry_is_tombstone(uint32_t e) { return (e & 0xFu) == ENTRY_TOMBSTONE; }

// lower bound by position in a sorted entry list whose positions are spread uniformly over [0, L) (mutation sites are
// uniform draws).  On the device: the interpolated guess is off by a few entries at most (binomial spread), so the eight
// entries around it are fetched with two 16-byte loads issued together (aligned on the arena, words outside the list are
// masked) and almost every search ends there after ONE memory latency; otherwise interpolation steps alternate with
// bisection steps on the rest (O(log n) on any list).  The arena is padded, so the window may overhang the list.
__host__ __device__ inline uint32_t entries_lower_interp(const uint32_t *e, uint32_t n, uint32_t pos, uint32_t L) {
    uint32_t lo = 0, hi = n, plo = 0, phi = L;  // positions of e[lo..hi) lie in [plo, phi]
    if (n == 0) return 0;
#ifdef __CUDA_ARCH__
    {
        uint32_t g = (uint32_t)((float)pos * ((float)n / (float)L));
        if (g >= n) g = n - 1;
        // window of W words starting at a 16-byte boundary of the arena, centred on the guess: 8 words for short lists,
        // 16 for long ones (the guess is off by ~sqrt(n) / 2 entries)
        const bool wide = n > 48;
        const uint32_t back = wide ? 6u : 2u;
        const uintptr_t first = (uintptr_t)(e + (g >= back ? g - back : 0)) & ~(uintptr_t)15;
        const int32_t k0 = (int32_t)(((intptr_t)first - (intptr_t)(uintptr_t)e) / 4);  // list index of the first window word (may be negative)
        const uint4 v0 = *reinterpret_cast<const uint4 *>(first), v1 = *reinterpret_cast<const uint4 *>(first + 16);
        uint4 v2 = make_uint4(0u, 0u, 0u, 0u), v3 = v2;
        if (wide) { v2 = *reinterpret_cast<const uint4 *>(first + 32); v3 = *reinterpret_cast<const uint4 *>(first + 48); }
        const uint32_t wv[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
        const int32_t W = wide ? 16 : 8;
        uint32_t below = 0, inside = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int32_t k = k0 + j;
            const bool in = j < W && k >= 0 && k < (int32_t)n;
            inside += in ? 1u : 0u;
            below += (in && entry_pos(wv[j]) < pos) ? 1u : 0u;
        }
        const int32_t kb = k0 < 0 ? 0 : k0;                     // first list index inside the window
        const int32_t ke = kb + (int32_t)inside;               // one past the last
        // the entries are sorted: `below` of them, the first ones of the window, are smaller
        if ((below > 0 || kb == 0) && (below < inside || ke == (int32_t)n)) return (uint32_t)kb + below;
        if (below == 0) { hi = (uint32_t)kb; } else { lo = (uint32_t)ke; }
    }
#endif
    bool interp = true;
    while (lo < hi) {
        uint32_t mid;
        if (interp && pos >= plo && phi > plo) {
            const float t = (float)(pos - plo) * ((float)(hi - lo) / (float)(phi - plo + 1u));
            const uint32_t ti = (uint32_t)t;
            mid = lo + (ti < hi - lo - 1u ? ti : hi - lo - 1u);
        } else {
            mid = (lo + hi) >> 1;
        }
        interp = !interp;
        const uint32_t p = entry_pos(e[mid]);
        if (p < pos) { lo = mid + 1; plo = p + 1u; } else { hi = mid; phi = p; }
    }
    return lo;
}

// merged view of a record: base list overridden by the delta, tombstones dropped; f(entry) in ascending position order
template <typename F>
__host__ __device__ inline void merged_for_each(const uint32_t *base, uint32_t len, const uint32_t *d, uint32_t nd, F &&f) {
    uint32_t a = 0, b = 0;
    while (a < len || b < nd) {
        const uint32_t pa = a < len ? entry_pos(base[a]) : 0xFFFFFFFFu, pb = b < nd ? entry_pos(d[b]) : 0xFFFFFFFFu;
        if (pa < pb) {
            f(base[a++]);
        } else {
            if (pa == pb) a++;
            const uint32_t e = d[b++];
            if (!entry_is_tombstone(e)) f(e);
        }
    }
}
// entry of the merged record at `pos`, or 0xFFFFFFFF when it has none
__host__ __device__ inline uint32_t merged_find(const uint32_t *base, uint32_t len, const uint32_t *d, uint32_t nd, uint32_t pos, uint32_t L) {
    for (uint32_t k = 0; k < nd; k++)
        if (entry_pos(d[k]) == pos) return entry_is_tombstone(d[k]) ? 0xFFFFFFFFu : d[k];
    const uint32_t i = entries_lower_interp(base, len, pos, L);
    if (i < len && entry_pos(base[i]) == pos) return base[i];
    return 0xFFFFFFFFu;
}

// lower bound by position in a sorted entry list
__host__ __device__ inline uint32_t entries_lower(const uint32_t *e, uint32_t n, uint32_t pos) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (entry_pos(e[mid]) < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// Haplotype::get_base (src/core/haplotype.rs:402-408)
__host__ __device__ inline uint32_t entries_get_base(const uint32_t *e, uint32_t n, uint32_t pos, const uint8_t *wt) {
    uint32_t i = entries_lower(e, n, pos);
    if (i < n && entry_pos(e[i]) == pos) return entry_g(e[i]);
    return wt[pos];
}
// get_mutations().get(pos): base index or 4 when absent
__host__ __device__ inline uint32_t entries_get_mut(const uint32_t *e, uint32_t n, uint32_t pos) {
    uint32_t i = entries_lower(e, n, pos);
    if (i < n && entry_pos(e[i]) == pos) return entry_m(e[i]);
    return 4;
}

// UtilityFunction::apply (src/core/fitness/utility.rs:48-60); same operation order as written there
__host__ __device__ inline double utility_apply(const DevProvider &p, double f) {
    if (p.util_kind == VL_UTILITY_ALGEBRAIC) {
        // let factor = 1. / (upper - 1.); upper * factor * f / (1. + factor * f): the two leading terms are constants
        return p.alg_uf * f / (1. + p.alg_factor * f);
    }
    if (p.util_kind == VL_UTILITY_HILL) {
        double fp = pow(f, p.hill_n);
        return fp / (p.hill_k + fp);
    }
    return f;
}

// nested-map row lookup: [lo,hi) of directed entries keyed k1
__host__ __device__ inline void epi_row(const DevProvider &p, uint32_t k1, uint32_t &lo_out, uint32_t &hi_out) {
#ifdef __CUDA_ARCH__
    if (p.epi_start && k1 < p.n_keys) {  // two loads instead of two binary searches over the whole table
        lo_out = p.epi_start[k1];
        hi_out = p.epi_start[k1 + 1];
        return;
    }
#endif
    uint32_t lo = 0, hi = p.n_epi;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (p.epi[mid].k1 < k1) lo = mid + 1; else hi = mid; }
    uint32_t b = lo;
    hi = p.n_epi;
    uint32_t a = lo;
    while (a < hi) { uint32_t mid = (a + hi) >> 1; if (p.epi[mid].k1 <= k1) a = mid + 1; else hi = mid; }
    lo_out = b; hi_out = a;
}
__host__ __device__ inline bool epi_get(const DevProvider &p, uint32_t lo, uint32_t hi, uint32_t k2, double &v) {
    uint32_t a = lo, b = hi;
    while (a < b) { uint32_t mid = (a + b) >> 1; if (p.epi[mid].k2 < k2) a = mid + 1; else b = mid; }
    if (a < hi && p.epi[a].k2 == k2) { v = p.epi[a].v; return true; }
    return false;
}
__host__ __device__ inline uint32_t ekey(uint32_t pos, uint32_t base) { return (pos << 2) | base; }

// FitnessTable::compute_fitness x EpistasisTable::compute_fitness over a flattened record
// (src/core/fitness/table.rs:72-79, src/core/fitness/epistasis.rs:176-193).  The reference multiplies in
// HashMap iteration order; here the order is ascending position (1e-12 relative tolerance applies, see
// DESIGN.md).  The epistatic factor keeps the reference's "partner absent" rule.
__host__ __device__ inline double compute_fitness_full(const DevProvider &p, const uint32_t *e, uint32_t n) {
    if (p.kind == VL_FITNESS_NEUTRAL) return 1.0;
    double acc = 1.;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t m = entry_m(e[k]);
        if (m < 4) acc = acc * p.table[(uint64_t)entry_pos(e[k]) * 4 + m];
    }
    if (p.kind == VL_FITNESS_EPISTATIC) {
        double fitness = 1.;
        for (uint32_t k = 0; k < n; k++) {
            uint32_t m = entry_m(e[k]);
            if (m >= 4) continue;
            uint32_t lo, hi;
            epi_row(p, ekey(entry_pos(e[k]), m), lo, hi);
            for (uint32_t q = lo; q < hi; q++) {
                uint32_t qpos = p.epi[q].k2 >> 2, qb = p.epi[q].k2 & 3;
                if (qpos > entry_pos(e[k]) && entries_get_mut(e, n, qpos) != qb) fitness *= p.epi[q].v;
            }
        }
        acc = acc * fitness;
    }
    return acc;
}

}  // namespace vl
