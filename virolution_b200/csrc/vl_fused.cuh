// vl_fused.cuh — the device-resident generation (vl_run, vl_next_generation) for compartments with uniform infection
// (every draw infects, hosts iid uniform: DevState::fast_infect), a reproductive fitness that does not depend on the
// host, no recombination and Binomial(L, mu) in the inversion regime — K1, K2 and K4 of SURVEY.md §8.
//
// The reference builds a host map (counting sort by host, src/core/hosts.rs:166-192) because BasicHost::replicate needs
// the fitness sum of every host's infectants (src/simulation.rs:137-152).  With H ~ N the slices have one member on
// average, so grouping costs far more than the sums it enables: a scatter of one record per infectant is bound by the
// SM's load/store pipeline (one 32-byte request per lane), and the measured sort + replicate kernels of that design spent
// their issue slots on slice bookkeeping (profiles/r2_a_*).  Here nothing is grouped:
//
//   k_infect_sum      host ~ U[0, H) and the Binomial(L, mu) count of every infectant from one Philox block; the host is
//                     kept per infectant (the reference's `infections`), the infectant's MAPPED FITNESS — carried with it
//                     since it was drawn — is added to hsum[host] by an f64 atomic; mutating infectants go to the worklist
//   k_mutate          (vl_kernels.cuh) new records; the new haplotype's mapped fitness replaces the carried value and is
//                     added to hsum[host]
//   k_offspring_sum   per infectant: S = hsum[host], lambda = f * R0 / S, Poisson(lambda) from the infectant's own
//                     stream; offspring counts stay in infectant order; clears the other hsum buffer for the next generation
//   k_plan_resample   (vl_kernels.cuh) sample size, exact multinomial split of the draws over tiles of RESAMPLE_TILE
//                     consecutive infectants
//   k_resample_pos    draws inside every tile; the drawn parent's haplotype AND mapped fitness become the child's
//
// Per infectant and generation: one scattered f64 reduction and one scattered 8-byte load on an H-entry array that L2
// holds (80 MB at H = 10^7), everything else streams.  No per-haplotype array is read except by k_mutate.
//
// Per-host sums: the reference adds a host's fitness values in slice order (src/simulation.rs:144).  An atomic sum adds
// them in arrival order.  For hosts with one or two members (74 % of the infectants at N = H) the two are the same f64
// (one addition, commutative); with three or more members they can differ in the last bit, which moves lambda by one ulp
// and changes a Poisson draw only if its uniform lies within ~1e-16 of a CDF value.  The replay path (bit-exactness
// against the recorded draws) does not use this file; tests/test_gpu_fused_scale.py checks that the populations of vl_run and
// of the trait-level calls agree infectant by infectant and that lambda agrees to 1e-15.
//
// Same draws as the trait-level path: hosts, mutation counts, mutations, offspring counts and resample draws are
// addressed by (seed, compartment, generation, infectant / output slot), and every path resamples over tiles of
// consecutive infectants (canonical_layout in virolution_gpu.cu).
#pragma once
#include "vl_kernels.cuh"

namespace vl {

// generation += 1 and the counters of a generation (one launch).  prefilled: the previous generation's resample has
// already drawn this generation's hosts and mutation counts (k_resample_pos<true>): the worklist is kept.
__device__ __forceinline__ void begin_generation_fused(DevState *st, int prefilled) {
    st->generation += 1;  // Simulation::increment_generation (src/simulation.rs:178)
    if (!prefilled) st->n_mut_list = 0;
    st->n_heavy = 0;
    st->n_medium = 0;
    st->fused_failed = 0;
    st->infectant_generations += st->n;
}
__global__ void k_begin_generation_fused(DevState *st, int prefilled) { begin_generation_fused(st, prefilled); }
// set_population of the device-resident generation: both per-infectant arrays swap, the sums change sides
__device__ __forceinline__ void swap_population_fused(DevState *st) {
    uint32_t *t = st->hap_id;
    st->hap_id = st->hap_id_next;
    st->hap_id_next = t;
    double *f = st->pop_fit;
    st->pop_fit = st->pop_fit_next;
    st->pop_fit_next = f;
    st->n = st->n_next;
    st->n_sorted = 0;
    st->hsum_parity ^= 1u;
}
__global__ void __launch_bounds__(TPB) k_zero_f64(double *p, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = 0.0;
}

// ---- infect + mutation count + per-host sum ---------------------------------------------------------------------------
// BasicSimulation::infect for uniform infection (src/simulation.rs:365-387 with one always-accepting draw per
// infectant) and the Binomial(L, mu) count of BasicHost::mutate (:94-95), both from block 0 of the infectant's PH_INFECT
// stream like k_infect / k_mutcount_fast.  The draws of infectant i in generation g depend on (seed, compartment, g, i)
// only, so they can be made as soon as i's haplotype is known: by k_infect_sum at the start of generation g, or by the
// resample of generation g - 1 for the children it has just drawn (k_resample_pos<true>) — same results either way.
struct InfectCtx {
    uint32_t H32, thr, gen;
    PhiloxKey key;
    uint32_t *host_id;
    double *hsum;
    uint64_t keep, stream;  // L2 policies: per-host sums stay, per-infectant streams pass
};
__device__ __forceinline__ void infect_draw(DevState *st, const InfectCtx &c, uint64_t i, double fit, uint32_t &h, uint32_t &nm) {
    const uint4 b0 = philox_block(c.key, i, c.gen, PH_INFECT, 0);
    uint64_t m = (uint64_t)b0.x * c.H32;  // Uniform(0, H): widening multiply, biased zone rejected exactly (as k_infect)
    if ((uint32_t)m < c.H32) {
        for (uint32_t blk = 1; (uint32_t)m < c.thr; blk++) m = (uint64_t)philox_block(c.key, i, c.gen, PH_INFECT, blk).x * c.H32;
    }
    h = (uint32_t)(m >> 32);
    nm = mutation_count_small(st, c.key, i, c.gen, b0);
    st_u32_hint(&c.host_id[i], h, c.stream);
    // (a mutating infectant's fitness is about to change: k_mutate stores the new value and adds it to the host's sum;
    // f == 0.0 adds nothing)
    if (nm == 0u && fit != 0.0) red_add_f64_hint(&c.hsum[h], fit, c.keep);
}
// Mutation worklist staged per CTA in shared memory: (infectant, n_mut, host, parent haplotype), appended with one vote
// and one shared atomic per warp round, flushed with ONE global atomic per CTA pass (same-address global atomics are
// served one after the other by L2).
constexpr uint32_t CL_CAP = RESAMPLE_TILE + RESAMPLE_TILE / 4;  // items the resample can stage between flushes (a tile's share of the draws is 1024 on average)
struct CtaList {
    uint4 *buf;
    uint32_t *cnt;  // [0] staged items, [1] global base of the flush
    uint32_t cap;   // entries of buf (the owner flushes before more than that can have been pushed)
};
__device__ __forceinline__ void cl_push(CtaList &cl, bool has, uint4 item) {  // (called by converged warps)
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, has);
    if (!mask) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&cl.cnt[0], (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (has) {
        const uint32_t k = base + __popc(mask & ((1u << lane) - 1u));
        if (k < cl.cap) cl.buf[k] = item;
    }
}
// (called by every thread of the CTA, after the pass's pushes) hap_out != nullptr: the id of a new record is the record
// count k_mutate starts with plus the item's worklist position — if k_mutate is the next kernel to touch the table, the
// infectant's haplotype can be written here, in stream order, instead of being patched by k_mutate
__device__ __forceinline__ void cl_flush(DevState *st, CtaList &cl, uint32_t *hap_out, unsigned long long id_base) {
    __syncthreads();
    const uint32_t n = cl.cnt[0] < cl.cap ? cl.cnt[0] : cl.cap;
    if (threadIdx.x == 0) {
        if (cl.cnt[0] > cl.cap) raise(st, DE_TILE_OVERFLOW);  // (cannot happen: every owner flushes in time)
        cl.cnt[1] = n ? atomicAdd(&st->n_mut_list, n) : 0u;
    }
    __syncthreads();
    const uint32_t gbase = cl.cnt[1];
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        const uint4 it = cl.buf[k];
        st->mut_list[gbase + k] = it;
        if (hap_out) hap_out[it.x] = (uint32_t)(id_base + gbase + k);
    }
    __syncthreads();
    if (threadIdx.x == 0) cl.cnt[0] = 0u;
    __syncthreads();
}
// GATHER: the mapped fitness comes from the per-haplotype memo (population installed by anything but k_resample_pos)
// and is stored per infectant, otherwise it came with the infectant.
constexpr int IS_TPB = 256;
constexpr int IS_ITEMS = RESAMPLE_TILE / IS_TPB;
template <bool GATHER>
__global__ void __launch_bounds__(IS_TPB, 4) k_infect_sum(DevState *st) {
    __shared__ uint4 s_items[RESAMPLE_TILE];  // one item per infectant of a chunk at most
    __shared__ uint32_t s_cnt[2];
    const uint64_t n = st->n;
    const uint32_t H32 = (uint32_t)st->H;
    if (H32 == 0) return;
    const uint32_t tid = threadIdx.x;
    InfectCtx c;
    c.H32 = H32;
    c.thr = (uint32_t)(-H32) % H32;
    c.gen = st->generation;
    c.key = philox_key(st->seed, st->compartment);
    c.host_id = st->host_id;
    c.hsum = st->hsum[st->hsum_parity];
    c.keep = l2_keep_policy();
    c.stream = l2_stream_policy();
    const unsigned long long id_base = st->n_hap;
    CtaList cl{s_items, s_cnt, RESAMPLE_TILE};
    if (tid == 0) s_cnt[0] = 0u;
    __syncthreads();
    uint32_t *hap_id = st->hap_id;
    double *pop_fit = st->pop_fit;
    const int pi0 = st->specs[0].repro;
    const double *__restrict__ hfit = pi0 >= 0 ? st->h_fit[pi0] : nullptr;
    const uint64_t n_chunks = (n + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
    for (uint64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        uint32_t hp[IS_ITEMS];
        double ft[IS_ITEMS];
#pragma unroll
        for (int e = 0; e < IS_ITEMS; e++) {
            const uint64_t i = ch * RESAMPLE_TILE + (uint64_t)e * IS_TPB + tid;
            hp[e] = i < n ? hap_id[i] : 0u;
        }
#pragma unroll
        for (int e = 0; e < IS_ITEMS; e++) {
            const uint64_t i = ch * RESAMPLE_TILE + (uint64_t)e * IS_TPB + tid;
            ft[e] = 0.0;
            if (i < n) {
                if (GATHER) { ft[e] = hfit ? hfit[hp[e]] : 1.0; pop_fit[i] = ft[e]; }
                else ft[e] = pop_fit[i];
            }
        }
#pragma unroll 2
        for (int e = 0; e < IS_ITEMS; e++) {
            const uint64_t i = ch * RESAMPLE_TILE + (uint64_t)e * IS_TPB + tid;
            uint32_t h = 0, nm = 0;
            if (i < n) infect_draw(st, c, i, ft[e], h, nm);
            cl_push(cl, nm != 0u, make_uint4((uint32_t)i, nm, h, hp[e]));
        }
        cl_flush(st, cl, hap_id, id_base);
    }
}

// ---- offspring counts in infectant order -------------------------------------------------------------------------------
// BasicHost::replicate (src/simulation.rs:120-154) per infectant: lambda_i = f_i * R0 / S_host, Poisson(lambda_i) from the
// infectant's PH_OFFSPRING stream (the same address k_offspring / k_replicate_fused use).  One CTA per resample tile
// (RESAMPLE_TILE consecutive infectants) so that the tile's offspring total comes out of the same pass.  The grid also
// clears the other hsum buffer: the next generation's sums start from zero without a pass of their own.
constexpr int OS_TPB = 256;
constexpr int OS_PER = RESAMPLE_TILE / OS_TPB;
template <bool SMALL_ONLY>
__global__ void __launch_bounds__(OS_TPB, 4) k_offspring_sum(DevState *st) {
    __shared__ uint64_t s_red[2][OS_TPB / 32];
    uint32_t it = 0;
    const uint64_t n = st->n;
    const uint64_t n_tiles = (n + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
    const uint32_t *__restrict__ host_id = st->host_id;
    const double *__restrict__ pop_fit = st->pop_fit;
    const double *hsum = st->hsum[st->hsum_parity];
    uint32_t *__restrict__ offspring = st->offspring;
    const uint32_t gen = st->generation;
    const double R0 = st->host_r0;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const uint32_t kmax = poisson_kmax(R0 < 12.0 ? R0 : 12.0);  // lambda = R0 * f / S <= R0 for non-negative fitness
    const uint32_t tid = threadIdx.x;
    const uint64_t keep = l2_keep_policy(), stream = l2_stream_policy();
    if (blockIdx.x == 0 && tid == 0) {
        st->n_sorted = n;
        st->host_tiles = 0;   // tiles of consecutive infectants (tile_range)
        st->n_mut_list = 0;   // k_mutate is done with this generation's worklist: the resample may start the next one
    }
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t p0 = tile * RESAMPLE_TILE;
        uint32_t h[OS_PER];
        double f[OS_PER], S[OS_PER];
#pragma unroll
        for (int j = 0; j < OS_PER; j++) {
            const uint64_t p = p0 + (uint64_t)j * OS_TPB + tid;
            h[j] = p < n ? ld_u32_hint(&host_id[p], stream) : NONE32;
            f[j] = p < n ? ld_f64_stream(&pop_fit[p], stream) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < OS_PER; j++) S[j] = h[j] != NONE32 ? ld_f64_hint(&hsum[h[j]], keep) : 1.0;
        uint32_t local = 0;
#pragma unroll
        for (int j = 0; j < OS_PER; j++) {
            const uint64_t p = p0 + (uint64_t)j * OS_TPB + tid;
            if (p >= n) continue;
            uint32_t off = 0;
            const double l = checked_lambda(st, f[j], R0, S[j]);  // lambda_i = f_i * R0 / S_host (src/simulation.rs:148-150)
            if (l > 0.0) {
                uint64_t v;
                if (l < 12.0) v = poisson_draw_small_fast(key, p, gen, PH_OFFSPRING, l, c_rcp64, kmax);
                else if (SMALL_ONLY) { raise(st, DE_UNDEFINED_OFFSPRING); v = 0; }  // f > S: negative fitness values in the host
                else v = poisson_ptrs(key, p, gen, PH_OFFSPRING, l);
                if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0xFFFFFFFFull; }
                off = (uint32_t)v;
            }
            st_u32_hint(&offspring[p], off, stream);
            local += off;
        }
        uint64_t tot = local;
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, sh);
        // one barrier per tile: the warps' totals alternate between two rows of shared memory
        uint64_t *row = s_red[it & 1u];
        it++;
        if ((tid & 31) == 0) row[tid >> 5] = tot;
        __syncthreads();
        if (tid == 0) {
            uint64_t t = 0;
            for (int w = 0; w < OS_TPB / 32; w++) t += row[w];
            st->tile_sum[tile] = t;
        }
    }
    // the buffer the next generation accumulates into
    double *other = st->hsum[st->hsum_parity ^ 1u];
    const uint64_t H = st->H;
    for (uint64_t k = (uint64_t)blockIdx.x * OS_TPB + tid; k < H; k += (uint64_t)gridDim.x * OS_TPB) st_f64_hint(&other[k], 0.0, keep);
}

// ---- resample over tiles of consecutive infectants ---------------------------------------------------------------------
// k_resample (vl_kernels.cuh) for offspring counts kept in infectant order next to hap_id[] / pop_fit[]: same tiles, same
// record order inside a tile, same draw values and the same value -> record rule as k_resample on the identity layout of
// the same offspring map, so the drawn haplotypes are the same.  The staged payload is (haplotype, mapped fitness): the
// child inherits both.
constexpr int RT_TPB = 256;
constexpr int RT_PER = RESAMPLE_TILE / RT_TPB;  // stripes: item q = k * RT_TPB + tid (coalesced loads)
static_assert(RANK_WORDS % RT_TPB == 0, "every thread owns RANK_WORDS / RT_TPB consecutive mask words");
struct ResamplePosArgs {
    uint32_t n_parts;
    uint32_t salt[MAX_PARTS];
    uint32_t *dst[MAX_PARTS];    // detached id buffer of every part but the own one
    double *dst_fit[MAX_PARTS];  // mapped fitness of the drawn parents (may be nullptr)
    int own_part;                // the part that goes straight into the population being assembled (hap_id_next, pop_fit_next), or -1
    const uint64_t *own_off;     // device: first slot of that part in the new population (nullptr = 0)
    int predict_ids;             // PREFILL: the next kernel to touch the record table is k_mutate, so the ids of the staged
                                 // children's new records are known and go into the population here (else k_mutate patches it)
    int tail;                    // what the last CTA to leave does: bit 0 = install the new population (k_swap_population_fused),
                                 // bit 1 = begin the next generation, whose infection this pass has prepared (k_begin_generation_fused)
};
// PREFILL (only when the caller enqueues the next generation right behind this one): the children written to the
// population (the own part) get their hosts and mutation counts for the NEXT generation right here (infect_draw with generation + 1): host_id[], the other hsum buffer (cleared by
// k_offspring_sum of this generation) and the worklist are ready when the next generation begins, which then starts with
// k_mutate.  The children are as they would be without it; k_infect_sum makes the same draws from the same streams.
// MULTI = false: one part, the whole new population (vl_run); nothing about parts is computed or kept in registers.
// SMALL: every Poisson mean is below 12 (host_r0 < 12, as k_offspring_sum<true>), so a tile of 1024 infectants cannot
// reach the 32768 offspring beyond which draws go through the CDF + guide table: that path is compiled out (a tile that
// does reach it is reported, DE_OFFSPRING_RANGE).
template <bool PREFILL, bool MULTI, bool SMALL = false>
__global__ void __launch_bounds__(RT_TPB, 4) k_resample_pos(DevState *st, ResamplePosArgs ra) {
    extern __shared__ uint4 s_items_dyn[];      // PREFILL: CL_CAP staged worklist items
    __shared__ uint32_t s_clcnt[2];
    __shared__ uint32_t s_hap[RESAMPLE_TILE];   // compacted payloads of the records with offspring
    __shared__ double s_fit[RESAMPLE_TILE];
    __shared__ uint32_t s_cdf[RESAMPLE_TILE];   // compacted inclusive CDF (guide path only)
    __shared__ uint16_t s_guide[GUIDE_SIZE];
    __shared__ uint32_t s_mask[RANK_WORDS + 1];
    __shared__ uint16_t s_pre[RANK_WORDS];
    __shared__ uint32_t s_wbits[RT_TPB / 32];
    __shared__ uint64_t s_pcnt[MAX_PARTS], s_pbase[MAX_PARTS];
    __shared__ uint32_t s_tv[RT_PER * (RT_TPB / 32)], s_tc[RT_PER * (RT_TPB / 32)];  // per (stripe, warp): offspring / non-zero records
    __shared__ uint32_t s_total;
    const uint64_t n = st->n;
    const uint64_t nt = (n + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
    const uint32_t *__restrict__ offspring = st->offspring;
    const uint32_t *__restrict__ hap_id = st->hap_id;
    const double *__restrict__ pop_fit = st->pop_fit;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr uint32_t NW = RT_TPB / 32;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const uint32_t gen = st->generation;
    const uint64_t stride = st->tile_stride;
    const uint64_t l2s = l2_stream_policy();
    const uint64_t own_off = MULTI && ra.own_off ? *ra.own_off : 0ull;
    const uint32_t n_parts = MULTI ? ra.n_parts : 1u;
    InfectCtx ictx;
    CtaList cl{s_items_dyn, s_clcnt, CL_CAP};
    if (PREFILL) {
        ictx.H32 = (uint32_t)st->H;
        ictx.thr = ictx.H32 ? (uint32_t)(-ictx.H32) % ictx.H32 : 0u;
        ictx.gen = gen + 1;
        ictx.key = key;
        ictx.host_id = st->host_id;
        ictx.hsum = st->hsum[st->hsum_parity ^ 1u];
        ictx.keep = l2_keep_policy();
        ictx.stream = l2_stream_policy();
        if (tid == 0) s_clcnt[0] = 0u;
    }
    for (uint64_t tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        uint64_t any = 0;
        __syncthreads();  // (a tile without draws is left early: nobody may still be reading the shared arrays)
        if (n_parts == 1) {
            any = st->tile_cnt[tile];
        } else {
            if (tid < n_parts) {
                s_pcnt[tid] = st->tile_cnt[tid * stride + tile];
                s_pbase[tid] = st->tile_base[tid * stride + tile];
            }
            __syncthreads();
            for (uint32_t p = 0; p < n_parts; p++) any |= s_pcnt[p];
        }
        const uint64_t tsum = st->tile_sum[tile];
        const uint64_t tbase1 = st->tile_base[tile];  // (single part; asked for with the other descriptors)
        const uint64_t start = tile * RESAMPLE_TILE;
        const uint32_t tlen = (uint32_t)(n - start < RESAMPLE_TILE ? n - start : RESAMPLE_TILE);
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
            v[k] = q < tlen ? ld_u32_hint(&offspring[start + q], l2s) : 0u;
            hp[k] = q < tlen ? ld_u32_hint(&hap_id[start + q], l2s) : 0u;
            ft[k] = q < tlen ? ld_f64_stream(&pop_fit[start + q], l2s) : 0.0;
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
        if (SMALL && total > RESAMPLE_RANK_MAX) {  // (uniform)
            if (tid == 0) raise(st, DE_OFFSPRING_RANGE);
            continue;
        }
        const bool direct = SMALL ? true : total <= RESAMPLE_RANK_MAX;
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
        if (n_parts > 1) {
            for (uint32_t p = 0; p < n_parts; p++) {
                const uint64_t c = s_pcnt[p], b = s_pbase[p];
                blocks_total += c ? (uint32_t)(((b + c - 1) >> 2) - (b >> 2) + 1) : 0u;
            }
        } else {
            blocks_total = (uint32_t)(((tbase1 + any - 1) >> 2) - (tbase1 >> 2) + 1);
        }
        // Every warp makes every pass (a pass = RT_TPB blocks = at most 1024 draws): the worklist votes need converged
        // warps.  Every child may mutate (mu * L large), so a tile with more draws than the list holds — far above its
        // average share — flushes between passes; the decision is the same for every thread of the CTA.
        const uint32_t passes = (blocks_total + RT_TPB - 1) / RT_TPB;
        const bool tight = PREFILL && 4ull * blocks_total > CL_CAP;
        for (uint32_t pass = 0; pass < passes; pass++) {
            if (tight && pass) cl_flush(st, cl, ra.predict_ids ? st->hap_id_next : nullptr, st->n_hap);
            const uint32_t e0 = pass * RT_TPB + wid * 32;
            const uint32_t e = e0 + lane;
            const bool live = e < blocks_total;
            uint32_t part = 0, before = 0;
            uint64_t c = any, base = tbase1;
            if (live && n_parts > 1) {
                for (;; part++) {
                    c = s_pcnt[part];
                    base = s_pbase[part];
                    const uint32_t nb = c ? (uint32_t)(((base + c - 1) >> 2) - (base >> 2) + 1) : 0u;
                    if (e < before + nb) break;
                    before += nb;
                }
            }
            const uint64_t salt = (uint64_t)ra.salt[part] << 48;
            const bool own = !MULTI || (int)part == ra.own_part;
            uint32_t *__restrict__ dst = own ? st->hap_id_next + own_off : ra.dst[part];
            double *__restrict__ dstf = own ? st->pop_fit_next + own_off : ra.dst_fit[part];
            const uint64_t q = (base >> 2) + (e - before);
            uint4 blk = make_uint4(0u, 0u, 0u, 0u);
            if (live) blk = philox_block(key, q | salt, gen, PH_RESAMPLE, 0);
            const uint32_t xs[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const uint64_t s = q * 4 + w;
                const bool valid = live && s >= base && s < base + c;
                uint32_t hap = 0, h = 0, nm = 0;
                if (valid) {
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
                    hap = s_hap[slot];
                    const double fit = s_fit[slot];
                    st_u32_hint(&dst[s], hap, l2s);
                    if (dstf) st_f64_hint(&dstf[s], fit, l2s);
                    if (PREFILL && own) infect_draw(st, ictx, own_off + s, fit, h, nm);
                }
                if (PREFILL) cl_push(cl, nm != 0u, make_uint4((uint32_t)(own_off + s), nm, h, hap));
            }
        }
        // (the staged children mutate at the start of the next generation, which follows at once (vl_run): their new
        // records' ids are known — record count + worklist position — and go into the population here.  Flushing after every
        // tile keeps the worklist in tile order; holding the items back until the list is half full was measured: slower,
        // 0.97 against 0.93 ms per generation, the mutate kernel included)
        if (PREFILL) cl_flush(st, cl, ra.predict_ids ? st->hap_id_next : nullptr, st->n_hap);
    }
    // The last CTA to leave finishes the generation (two one-thread launches less per generation): every other CTA's
    // writes are behind its fence, nobody reads the pointers any more.
    if (ra.tail) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(&st->tail_done, 1u) == gridDim.x - 1) {
                st->tail_done = 0u;
                if (ra.tail & 1) swap_population_fused(st);
                if (ra.tail & 2) begin_generation_fused(st, 1);
            }
        }
    }
}

// set_population of the device-resident generation: both per-infectant arrays swap, the sums change sides
__global__ void k_swap_population_fused(DevState *st) { swap_population_fused(st); }

}  // namespace vl
