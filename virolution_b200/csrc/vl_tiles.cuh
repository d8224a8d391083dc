// vl_tiles.cuh — the tile-strided generation: the device-resident loop (vl_run, vl_next_generation) for compartments with
// uniform infection (every draw infects, hosts iid uniform: DevState::fast_infect), one reproductive provider, no
// recombination and Binomial(L, mu) in the inversion regime — K1, K2 and K4 of SURVEY.md §8.
//
// What the reference does per generation (src/simulation.rs:200-218): infect (:365-387), build the host map
// (src/core/hosts.rs:166-192), mutate (:389-428), replicate (:430-485), subsample (:516-527).  Here, per generation:
//
//   k_prefill          infect + Binomial(L, mu) count per infectant, one Philox block each; the record (infectant,
//                      haplotype, MAPPED FITNESS, host-in-tile) is appended to the region of its host tile (a power-of-two
//                      range of hosts), mutating infectants reserve their slot and go to the worklist
//   k_mutate           (vl_kernels.cuh) writes the records of the mutated infectants into the reserved slots
//   k_sort_replicate   one CTA per host tile: counting sort by host in shared memory, the reference's descending in-host
//                      order, per-host fitness sum in that order, lambda, Poisson — all from the staged tile; writes
//                      (offspring, haplotype, mapped fitness) per tile position
//   k_plan_resample    (vl_kernels.cuh) sample size and exact multinomial split of the draws over tiles
//   k_resample_tl      draws inside every tile; the drawn parent's haplotype AND mapped fitness become the child's
//
// The mapped fitness travels with the infectant (16-byte records) instead of being gathered from the per-haplotype memo
// once per infectant and generation: that gather was one random 32-byte sector per infectant over a table of 10^7..10^8
// records (L2 hit rate 29 %, 3.4x the algorithmic traffic of the old replicate kernel).  Nothing in this flow reads a
// per-haplotype array except k_mutate (one 64-byte line per NEW record).  There are no dense CSR positions and no
// offsets[] array: tile t owns [t * TL_CAP, t * TL_CAP + fill[t]).
//
// Same draws as the trait-level path: hosts, mutation counts, offspring counts and resample draws are addressed by
// (seed, compartment, generation, infectant / output slot), tiles are the same host ranges (host_tile_size), records
// inside a tile are in the same order (host ascending, infectant descending) — so vl_run(k) and k generations of
// trait-level calls give the same population (tests/test_gpu_stochastic.py, tests/test_gpu_fused_scale.py).
#pragma once
#include "vl_kernels.cuh"

namespace vl {

// generation += 1, counters cleared, tile geometry of this generation decided from the population size on the device
__global__ void __launch_bounds__(1024) k_begin_generation_tl(DevState *st) {
    const uint64_t n = st->n;
    const uint32_t ht = host_tile_size(n, st->H);
    const uint64_t nt = ht ? (st->H + ht - 1) / ht : 0;
    const uint64_t cap = st->tl_tiles_cap;
    for (uint64_t t = threadIdx.x; t < nt && t < cap; t += blockDim.x) st->tl_fill[t * TL_FILL_STRIDE] = 0u;
    if (threadIdx.x == 0) {
        st->generation += 1;  // Simulation::increment_generation (src/simulation.rs:178)
        st->n_mut_list = 0;
        st->n_heavy = 0;
        st->n_medium = 0;
        st->fused_failed = 0;
        st->ht_planned = ht;
        st->tl_shift = ht ? 31u - (uint32_t)__clz(ht) : 0u;
        st->infectant_generations += n;
        if (n > 0 && (ht == 0 || nt > cap)) raise(st, DE_TILE_OVERFLOW);
    }
}

// ---- infect + mutation count + append to the host tile ---------------------------------------------------------------
// BasicSimulation::infect for uniform infection (src/simulation.rs:365-387 with one always-accepting draw per
// infectant) and the Binomial(L, mu) count of BasicHost::mutate (:94-95), both from block 0 of the infectant's PH_INFECT
// stream like k_infect / k_mutcount_fast.  GATHER: the mapped fitness comes from the per-haplotype memo (population
// installed by anything but k_resample_tl), otherwise from pop_fit[] (it came with the infectant).
constexpr int PF_TPB = 256;
constexpr int PF_ITEMS = 4;
template <bool GATHER>
__global__ void __launch_bounds__(PF_TPB, 4) k_prefill(DevState *st) {
    __shared__ uint32_t s_base;
    __shared__ uint32_t s_w[PF_TPB / 32];
    const uint64_t n = st->n;
    const uint32_t H32 = (uint32_t)st->H;
    const uint32_t gen = st->generation, shift = st->tl_shift, hmask = (1u << shift) - 1u;
    const uint64_t tiles_cap = st->tl_tiles_cap;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const uint32_t *__restrict__ hap_id = st->hap_id;
    const double *__restrict__ pop_fit = st->pop_fit;
    const int pi0 = st->specs[0].repro;
    const double *__restrict__ hfit = pi0 >= 0 ? st->h_fit[pi0] : nullptr;
    uint32_t *fill = st->tl_fill;
    uint4 *__restrict__ tl_rec = st->tl_rec;
    uint16_t *__restrict__ tl_hl = st->tl_hl;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (H32 == 0 || st->ht_planned == 0) return;
    const uint32_t thr = (uint32_t)(-H32) % H32;
    constexpr uint64_t CHUNK = (uint64_t)PF_TPB * PF_ITEMS;
    const uint64_t n_chunks = (n + CHUNK - 1) / CHUNK;
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        uint32_t hp[PF_ITEMS], hn[PF_ITEMS], pos[PF_ITEMS];  // hn: host-in-tile | n_mut << 16, pos: tile, then slot
        double ft[PF_ITEMS];
#pragma unroll
        for (int e = 0; e < PF_ITEMS; e++) {
            const uint64_t i = c * CHUNK + (uint64_t)e * PF_TPB + tid;
            hp[e] = i < n ? hap_id[i] : 0u;
        }
#pragma unroll
        for (int e = 0; e < PF_ITEMS; e++) {
            const uint64_t i = c * CHUNK + (uint64_t)e * PF_TPB + tid;
            ft[e] = 1.0;
            if (i < n) ft[e] = GATHER ? (hfit ? hfit[hp[e]] : 1.0) : pop_fit[i];
        }
#pragma unroll
        for (int e = 0; e < PF_ITEMS; e++) {
            const uint64_t i = c * CHUNK + (uint64_t)e * PF_TPB + tid;
            pos[e] = NONE32;
            hn[e] = 0;
            if (i >= n) continue;
            const uint4 b0 = philox_block(key, i, gen, PH_INFECT, 0);
            uint64_t m = (uint64_t)b0.x * H32;  // Uniform(0, H): widening multiply, biased zone rejected exactly (as k_infect)
            if ((uint32_t)m < H32) {
                for (uint32_t blk = 1; (uint32_t)m < thr; blk++) m = (uint64_t)philox_block(key, i, gen, PH_INFECT, blk).x * H32;
            }
            const uint32_t h = (uint32_t)(m >> 32);
            const uint32_t nm = min(mutation_count_small(st, key, i, gen, b0), 0xFFFFu);
            pos[e] = h >> shift;
            hn[e] = (h & hmask) | (nm << 16);
        }
#pragma unroll
        for (int e = 0; e < PF_ITEMS; e++) {  // one append cursor per tile; all of a thread's atomics are in flight together
            const uint32_t tile = pos[e];
            if (tile == NONE32) continue;
            if (tile >= tiles_cap) { raise(st, DE_TILE_OVERFLOW); pos[e] = NONE32; continue; }
            const uint32_t r = atomicAdd(&fill[(uint64_t)tile * TL_FILL_STRIDE], 1u);
            pos[e] = r < TL_CAP ? tile * TL_CAP + r : NONE32 - 1u;
        }
        uint32_t mine = 0;
#pragma unroll
        for (int e = 0; e < PF_ITEMS; e++) {
            const uint64_t i = c * CHUNK + (uint64_t)e * PF_TPB + tid;
            if (i >= n || pos[e] == NONE32) continue;
            if (pos[e] == NONE32 - 1u) { raise(st, DE_TILE_OVERFLOW); pos[e] = NONE32; }
            if ((hn[e] >> 16) == 0u) {
                if (pos[e] != NONE32) {
                    const unsigned long long fb = (unsigned long long)__double_as_longlong(ft[e]);
                    tl_rec[pos[e]] = make_uint4((uint32_t)i, hp[e], (uint32_t)fb, (uint32_t)(fb >> 32));
                    tl_hl[pos[e]] = (uint16_t)(hn[e] & 0xFFFFu);
                }
            } else {
                mine += 1u;
            }
        }
        // mutation worklist: (infectant, n_mut | host-in-tile << 16, reserved slot, parent haplotype); one global atomic per chunk
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < PF_TPB / 32; w++) {
            const uint32_t t = s_w[w];
            if (w < (int)wid) woff += t;
            tot += t;
        }
        if (tid == 0) s_base = tot ? atomicAdd(&st->n_mut_list, tot) : 0u;
        __syncthreads();
        if (mine) {
            uint32_t o = s_base + woff + inc - mine;
#pragma unroll
            for (int e = 0; e < PF_ITEMS; e++) {
                const uint64_t i = c * CHUNK + (uint64_t)e * PF_TPB + tid;
                const uint32_t nm = hn[e] >> 16;
                if (i < n && nm) st->mut_list[o++] = make_uint4((uint32_t)i, nm | ((hn[e] & 0xFFFFu) << 16), pos[e], hp[e]);
            }
        }
    }
}

// ---- sort by host + replicate, one CTA per host tile -----------------------------------------------------------------
// HostMapBuffer::build (src/core/hosts.rs:166-192) restricted to the tile's hosts, then BasicHost::replicate
// (src/simulation.rs:120-154) for every host of the tile:
//   1. records from tl_rec into registers, histogram over the tile's hosts (shared atomics give each record an arrival
//      number inside its host),
//   2. exclusive scan of the histogram: the slice of every host,
//   3. infectant indices into the slices (arrival order), then every record ranks itself by DESCENDING infectant index
//      inside its slice — the order the reference's back-fill leaves (KAT src/core/hosts.rs:317-318) — and moves there,
//   4. every position sums its slice in slice order from 0.0, forms lambda = f * R0 / S, draws Poisson(lambda) from the
//      infectant's PH_OFFSPRING stream, and the tile's (offspring, haplotype, fitness) rows go out coalesced.
// Same arithmetic and the same Philox addresses as k_replicate_fused on the same host tile.
constexpr int SR_TPB = 512;
constexpr int SR_PER = TL_CAP / SR_TPB;
constexpr uint32_t SR_MAX_SLICE = 31;  // members of one host a tile handles (uniform infection: Poisson(<= 6.7) occupancy)
constexpr int SR_SMEM_BYTES = (HT_MAX_HOSTS + 4) * 4 + TL_CAP * (4 + 4 + 4 + 2 + 8);
static_assert(HT_MAX_HOSTS % SR_TPB == 0 && TL_CAP % SR_TPB == 0, "whole hosts / records per thread");
template <bool SMALL_ONLY>
__global__ void __launch_bounds__(SR_TPB, 3) k_sort_replicate(DevState *st) {
    extern __shared__ __align__(16) unsigned char s_dyn_sr[];
    double *s_ofit = reinterpret_cast<double *>(s_dyn_sr);                     // TL_CAP mapped fitness, slice order
    uint32_t *s_off = reinterpret_cast<uint32_t *>(s_ofit + TL_CAP);           // HT_MAX_HOSTS + 1 counts, then slice starts
    uint32_t *s_idx = s_off + HT_MAX_HOSTS + 4;                                // TL_CAP infectant indices, arrival order
    uint32_t *s_oidx = s_idx + TL_CAP;                                         // TL_CAP infectant indices, slice order
    uint32_t *s_ohap = s_oidx + TL_CAP;                                        // TL_CAP haplotypes, slice order
    uint16_t *s_seg = reinterpret_cast<uint16_t *>(s_ohap + TL_CAP);           // TL_CAP slice start | length << 11
    __shared__ uint32_t s_w[SR_TPB / 32];
    __shared__ uint64_t s_red64[SR_TPB / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t shift = st->tl_shift, ht = 1u << shift;
    const uint64_t H = st->H;
    if (st->ht_planned == 0) return;
    const uint64_t nt = (H + ht - 1) >> shift;
    const uint32_t gen = st->generation;
    const double R0 = st->host_r0;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const uint32_t kmax = poisson_kmax(R0 < 12.0 ? R0 : 12.0);  // lambda = R0 * f / S <= R0 for non-negative fitness
    const uint32_t *fill = st->tl_fill;
    const uint4 *__restrict__ tl_rec = st->tl_rec;
    const uint16_t *__restrict__ tl_hl = st->tl_hl;
    uint4 *__restrict__ out = st->tl_sorted;
    if (blockIdx.x == 0 && tid == 0) { st->host_tiles = ht; st->layout = 1u; st->n_sorted = st->n; }
    for (uint64_t tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        const uint64_t h0 = tile << shift;
        const uint32_t nh = (uint32_t)(H - h0 < ht ? H - h0 : ht);
        const uint32_t F = fill[tile * TL_FILL_STRIDE];
        const uint64_t base = tile * TL_CAP;
        if (F > TL_CAP) {  // (uniform across the CTA) reported by k_prefill already
            if (tid == 0) st->tile_sum[tile] = 0;
            continue;
        }
        uint4 r[SR_PER];
        uint32_t ha[SR_PER];  // host-in-tile | arrival << 16
#pragma unroll
        for (int j = 0; j < SR_PER; j++) {
            const uint32_t q = tid + j * SR_TPB;
            r[j] = q < F ? tl_rec[base + q] : make_uint4(0u, 0u, 0u, 0u);
            ha[j] = q < F ? (uint32_t)tl_hl[base + q] : 0u;
        }
        __syncthreads();  // the previous tile's shared arrays are free
        for (uint32_t k = tid; k <= nh; k += SR_TPB) s_off[k] = 0u;
        __syncthreads();
        bool failed = false;
#pragma unroll
        for (int j = 0; j < SR_PER; j++) {
            const uint32_t q = tid + j * SR_TPB;
            if (q >= F) continue;
            const uint32_t hl = ha[j] < nh ? ha[j] : 0u;
            if (ha[j] >= nh) failed = true;  // (cannot happen: the host came from a draw below H)
            const uint32_t arr = atomicAdd(&s_off[hl], 1u);
            if (arr >= SR_MAX_SLICE) failed = true;
            ha[j] = hl | (arr << 16);
        }
        if (__syncthreads_or(failed ? 1 : 0)) {
            if (tid == 0) { raise(st, DE_TILE_OVERFLOW); st->tile_sum[tile] = 0; }
            continue;
        }
        {   // exclusive scan of the counts: HT_MAX_HOSTS / SR_TPB consecutive hosts per thread
            constexpr uint32_t HP = HT_MAX_HOSTS / SR_TPB;
            uint32_t c[HP], sum = 0;
#pragma unroll
            for (uint32_t u = 0; u < HP; u++) { c[u] = tid * HP + u < nh ? s_off[tid * HP + u] : 0u; sum += c[u]; }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= (uint32_t)d) inc += y;
            }
            if (lane == 31) s_w[wid] = inc;
            __syncthreads();
            uint32_t run = inc - sum;
#pragma unroll
            for (int w = 0; w < SR_TPB / 32; w++)
                if (w < (int)wid) run += s_w[w];
#pragma unroll
            for (uint32_t u = 0; u < HP; u++) {
                if (tid * HP + u < nh) s_off[tid * HP + u] = run;
                run += c[u];
            }
            if (tid == 0) s_off[nh] = F;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SR_PER; j++) {
            const uint32_t q = tid + j * SR_TPB;
            if (q < F) s_idx[s_off[ha[j] & 0xFFFFu] + (ha[j] >> 16)] = r[j].x;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SR_PER; j++) {
            const uint32_t q = tid + j * SR_TPB;
            if (q >= F) continue;
            const uint32_t hl = ha[j] & 0xFFFFu;
            const uint32_t b = s_off[hl], e = s_off[hl + 1];
            uint32_t rank = 0;
            for (uint32_t a = b; a < e; a++) rank += s_idx[a] > r[j].x ? 1u : 0u;
            const uint32_t f = b + rank;
            s_oidx[f] = r[j].x;
            s_ohap[f] = r[j].y;
            s_ofit[f] = __longlong_as_double((long long)(((unsigned long long)r[j].w << 32) | r[j].z));
            s_seg[f] = (uint16_t)(b | ((e - b) << 11));
        }
        __syncthreads();
        uint32_t local = 0;
#pragma unroll
        for (int j = 0; j < SR_PER; j++) {
            const uint32_t q = tid + j * SR_TPB;
            if (q >= F) continue;
            const uint32_t sg = s_seg[q];
            const uint32_t b = sg & 0x7FFu, e = b + (sg >> 11);
            double S = 0.0;
            for (uint32_t a = b; a < e; a++) S += s_ofit[a];  // slice order from 0.0: iter().sum() (src/simulation.rs:144)
            const double fq = s_ofit[q];
            const double l = checked_lambda(st, fq, R0, S);   // lambda_i = f_i * R0 / S_host (src/simulation.rs:148-150)
            uint32_t off = 0;
            if (l > 0.0) {
                const uint64_t who = s_oidx[q];
                uint64_t v;
                if (l < 12.0) v = poisson_draw_small_fast(key, who, gen, PH_OFFSPRING, l, c_rcp64, kmax);
                else if (SMALL_ONLY) { raise(st, DE_UNDEFINED_OFFSPRING); v = 0; }  // f > S: negative fitness values in the slice
                else v = poisson_ptrs(key, who, gen, PH_OFFSPRING, l);
                if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0xFFFFFFFFull; }
                off = (uint32_t)v;
            }
            const unsigned long long fb = (unsigned long long)__double_as_longlong(fq);
            out[base + q] = make_uint4(off, s_ohap[q], (uint32_t)fb, (uint32_t)(fb >> 32));
            local += off;
        }
        uint64_t tot = local;
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, sh);
        if (lane == 0) s_red64[wid] = tot;
        __syncthreads();
        if (tid == 0) {
            uint64_t t = 0;
            for (int w = 0; w < SR_TPB / 32; w++) t += s_red64[w];
            st->tile_sum[tile] = t;
        }
    }
}

// ---- resample inside the tiles of the tile-strided layout ------------------------------------------------------------
// k_resample (vl_kernels.cuh) for DevState::layout == 1: same tiles, same record order inside a tile, same draw values and
// the same value -> record rule, so the drawn haplotypes are the ones k_resample draws from the dense layout of the same
// host map.  The staged payload is (haplotype, mapped fitness): the child inherits both (mode 0 only).
constexpr int RT_TPB = 512;
constexpr int RT_PER = TL_CAP / RT_TPB;  // stripes: item q = k * RT_TPB + tid (coalesced 16-byte loads)
static_assert(RANK_WORDS % RT_TPB == 0, "every thread owns RANK_WORDS / RT_TPB consecutive mask words");
struct ResampleTlArgs {
    uint32_t n_parts;
    uint32_t salt[MAX_PARTS];
    uint32_t *dst[MAX_PARTS];    // nullptr = the population being assembled (hap_id_next, pop_fit_next)
    double *dst_fit[MAX_PARTS];  // mapped fitness of the drawn parents (may be nullptr when dst is a detached buffer)
};
__global__ void __launch_bounds__(RT_TPB, 2) k_resample_tl(DevState *st, ResampleTlArgs ra) {
    __shared__ uint32_t s_hap[TL_CAP];   // compacted payloads of the records with offspring
    __shared__ double s_fit[TL_CAP];
    __shared__ uint32_t s_cdf[TL_CAP];   // compacted inclusive CDF (guide path only)
    __shared__ uint16_t s_guide[GUIDE_SIZE];
    __shared__ uint32_t s_mask[RANK_WORDS + 1];
    __shared__ uint16_t s_pre[RANK_WORDS];
    __shared__ uint32_t s_wbits[RT_TPB / 32];
    __shared__ uint64_t s_pcnt[MAX_PARTS], s_pbase[MAX_PARTS];
    __shared__ uint32_t s_tv[RT_PER * (RT_TPB / 32)], s_tc[RT_PER * (RT_TPB / 32)];  // per (stripe, warp): offspring / non-zero records
    __shared__ uint32_t s_total;
    const uint64_t nt = tile_count(st);
    const uint4 *__restrict__ rows = st->tl_sorted;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr uint32_t NW = RT_TPB / 32;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const uint32_t gen = st->generation;
    const uint64_t stride = st->tile_stride;
    for (uint64_t tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        uint64_t any = 0;
        __syncthreads();  // (a tile without draws is left early: nobody may still be reading the shared arrays)
        if (ra.n_parts == 1) {
            any = st->tile_cnt[tile];
        } else {
            if (tid < ra.n_parts) {
                s_pcnt[tid] = st->tile_cnt[tid * stride + tile];
                s_pbase[tid] = st->tile_base[tid * stride + tile];
            }
            __syncthreads();
            for (uint32_t p = 0; p < ra.n_parts; p++) any |= s_pcnt[p];
        }
        const uint64_t tsum = st->tile_sum[tile];
        uint64_t start;
        uint32_t tlen;
        tile_range(st, tile, start, tlen);
        if (any == 0) continue;  // uniform across the CTA
        if (tsum > 0xFFFFFFFFull) {  // the tile CDF is 32-bit
            if (tid == 0) raise(st, DE_OFFSPRING_RANGE);
            continue;
        }
        // ---- stripes of RT_TPB consecutive positions: warp scans of offspring and of the non-zero flags
        uint32_t v[RT_PER], hp[RT_PER], iv[RT_PER], ic[RT_PER];
        double ft[RT_PER];
#pragma unroll
        for (int k = 0; k < RT_PER; k++) {
            const uint32_t q = k * RT_TPB + tid;
            uint4 x = make_uint4(0u, 0u, 0u, 0u);
            if (q < tlen) x = rows[start + q];
            v[k] = x.x;
            hp[k] = x.y;
            ft[k] = __longlong_as_double((long long)(((unsigned long long)x.w << 32) | x.z));
        }
        constexpr uint32_t WPT = RANK_WORDS / RT_TPB;  // mask words per thread
#pragma unroll
        for (uint32_t k = 0; k < WPT; k++) s_mask[tid * WPT + k] = 0u;
        if (tid == 0) s_mask[RANK_WORDS] = 0u;
#pragma unroll
        for (int k = 0; k < RT_PER; k++) {
            uint32_t a = v[k], c = v[k] ? 1u : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, a, d), z = __shfl_up_sync(0xFFFFFFFFu, c, d);
                if (lane >= (uint32_t)d) { a += y; c += z; }
            }
            iv[k] = a;
            ic[k] = c;
            if (lane == 31) { s_tv[k * NW + wid] = a; s_tc[k * NW + wid] = c; }
        }
        __syncthreads();
        if (wid == 0) {  // exclusive scan of the RT_PER * NW (stripe, warp) totals: two entries per lane
            constexpr uint32_t NE = RT_PER * NW, EPL = (NE + 31) / 32;
            uint32_t xa[EPL], xc[EPL], sa = 0, sc = 0;
#pragma unroll
            for (uint32_t u = 0; u < EPL; u++) {
                const uint32_t idx = lane * EPL + u;
                xa[u] = idx < NE ? s_tv[idx] : 0u;
                xc[u] = idx < NE ? s_tc[idx] : 0u;
                sa += xa[u];
                sc += xc[u];
            }
            uint32_t ia = sa, icc = sc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, ia, d), z = __shfl_up_sync(0xFFFFFFFFu, icc, d);
                if (lane >= (uint32_t)d) { ia += y; icc += z; }
            }
            if (lane == 31) s_total = ia;
            uint32_t ra_ = ia - sa, rc_ = icc - sc;
#pragma unroll
            for (uint32_t u = 0; u < EPL; u++) {
                const uint32_t idx = lane * EPL + u;
                if (idx < NE) { s_tv[idx] = ra_; s_tc[idx] = rc_; }
                ra_ += xa[u];
                rc_ += xc[u];
            }
        }
        __syncthreads();
        const uint32_t total = s_total;
        const bool direct = total <= RESAMPLE_RANK_MAX;
        const double scale = (double)GUIDE_SIZE / (double)total;
#pragma unroll
        for (int k = 0; k < RT_PER; k++) {
            if (!v[k]) continue;
            const uint32_t lo = s_tv[k * NW + wid] + iv[k] - v[k];  // exclusive prefix of the offspring counts
            const uint32_t slot = s_tc[k * NW + wid] + ic[k] - 1u;  // compacted index
            s_hap[slot] = hp[k];
            s_fit[slot] = ft[k];
            if (direct) {
                atomicOr(&s_mask[lo >> 5], 1u << (lo & 31));
            } else {
                const uint32_t hi = lo + v[k];
                s_cdf[slot] = hi;
                // guide[b] = this entry for every bucket b whose first reachable draw value (b*total) >> GUIDE_BITS lies in [lo, hi)
                uint32_t bf = (uint32_t)((double)lo * scale);
                if (bf > GUIDE_SIZE) bf = GUIDE_SIZE;
                while (bf > 0 && (uint32_t)(((uint64_t)(bf - 1) * total) >> GUIDE_BITS) >= lo) bf--;
                while (bf < GUIDE_SIZE && (uint32_t)(((uint64_t)bf * total) >> GUIDE_BITS) < lo) bf++;
                for (uint32_t b = bf; b < GUIDE_SIZE && (uint32_t)(((uint64_t)b * total) >> GUIDE_BITS) < hi; b++) s_guide[b] = (uint16_t)slot;
            }
        }
        __syncthreads();
        if (direct) {  // set bits before every mask word: per-thread popc, warp scan, warp totals
            uint32_t bits[WPT], mine = 0;
#pragma unroll
            for (uint32_t k = 0; k < WPT; k++) { bits[k] = __popc(s_mask[tid * WPT + k]); mine += bits[k]; }
            uint32_t binc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, binc, d);
                if (lane >= (uint32_t)d) binc += y;
            }
            if (lane == 31) s_wbits[wid] = binc;
            __syncthreads();
            uint32_t before = binc - mine;
#pragma unroll
            for (int w = 0; w < (int)NW; w++)
                if (w < (int)wid) before += s_wbits[w];
#pragma unroll
            for (uint32_t k = 0; k < WPT; k++) { s_pre[tid * WPT + k] = (uint16_t)before; before += bits[k]; }
            __syncthreads();
        }
        // ---- draws of every part: output slots [base, base + c), four per Philox block (as k_resample)
        uint32_t blocks_total = 0;
        if (ra.n_parts > 1) {
            for (uint32_t p = 0; p < ra.n_parts; p++) {
                const uint64_t c = s_pcnt[p], b = s_pbase[p];
                blocks_total += c ? (uint32_t)(((b + c - 1) >> 2) - (b >> 2) + 1) : 0u;
            }
        } else {
            const uint64_t b = st->tile_base[tile];
            blocks_total = (uint32_t)(((b + any - 1) >> 2) - (b >> 2) + 1);
        }
        for (uint32_t e = tid; e < blocks_total; e += RT_TPB) {
            uint32_t part = 0, before = 0;
            uint64_t c = any, base = 0;
            if (ra.n_parts > 1) {
                for (;; part++) {
                    c = s_pcnt[part];
                    base = s_pbase[part];
                    const uint32_t nb = c ? (uint32_t)(((base + c - 1) >> 2) - (base >> 2) + 1) : 0u;
                    if (e < before + nb) break;
                    before += nb;
                }
            } else {
                base = st->tile_base[tile];
            }
            const uint64_t salt = (uint64_t)ra.salt[part] << 48;
            uint32_t *__restrict__ dst = ra.dst[part] ? ra.dst[part] : st->hap_id_next;
            double *__restrict__ dstf = ra.dst[part] ? ra.dst_fit[part] : st->pop_fit_next;
            const uint64_t q = (base >> 2) + (e - before);
            const uint4 blk = philox_block(key, q | salt, gen, PH_RESAMPLE, 0);
            const uint32_t xs[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const uint64_t s = q * 4 + w;
                if (s < base || s >= base + c) continue;
                uint32_t x = xs[w];
                uint64_t mm = (uint64_t)x * total;
                if ((uint32_t)mm < total) {  // biased zone of the widening multiply: exact rejection
                    const uint32_t thr = (uint32_t)(-total) % total;
                    if ((uint32_t)mm < thr) {
                        Philox g(st->seed, st->compartment, gen, PH_RESAMPLE_RETRY, s | salt);
                        do { x = g.u32(); mm = (uint64_t)x * total; } while ((uint32_t)mm < thr);
                    }
                }
                const uint32_t r = (uint32_t)(mm >> 32);
                uint32_t slot;
                if (direct) {
                    const uint32_t upto = s_mask[r >> 5] & (0xFFFFFFFFu >> (31 - (r & 31)));  // run starts at or below r in its word
                    slot = (uint32_t)s_pre[r >> 5] + __popc(upto) - 1u;
                } else {
                    slot = s_guide[x >> (32 - GUIDE_BITS)];
                    while (s_cdf[slot] <= r) slot++;
                }
                dst[s] = s_hap[slot];
                if (dstf) dstf[s] = s_fit[slot];
            }
        }
    }
}

// set_population of the tile-strided generation: both per-infectant arrays swap
__global__ void k_swap_population_tl(DevState *st) {
    uint32_t *t = st->hap_id;
    st->hap_id = st->hap_id_next;
    st->hap_id_next = t;
    double *f = st->pop_fit;
    st->pop_fit = st->pop_fit_next;
    st->pop_fit_next = f;
    st->n = st->n_next;
    st->n_sorted = 0;
    st->layout = 0u;
}

}  // namespace vl
