// vl_kernels.cuh — the per-generation population update as sm_100a kernels.
//
// Kernel                  reference rows it replaces
//   k_infect              BasicSimulation::infect + BasicHost::infect            src/simulation.rs:365-387,56-65
//                         (+ the Binomial(L, mu) mutation count of BasicHost::mutate, :94-95, fused: the count
//                         depends only on (seed, generation, infectant), not on the phase that draws it)
//   scan, k_scatter       HostMapBuffer::build (counting sort, back-fill)         src/core/hosts.rs:166-192
//   k_fitness, k_host_sum[_heavy]  the descending in-host order that back-fill produces, then
//                         BasicHost::replicate's fitness, per-host sum and lambda  src/simulation.rs:120-147
//   k_recombine           BasicHost::mutate, recombination part                   src/simulation.rs:68-90
//   k_mutation_counts,    BasicHost::mutate, mutation part + create_descendant    src/simulation.rs:93-117,
//   k_mutate                + FitnessProvider::get_fitness for every provider     src/core/haplotype.rs:241-260
//   k_offspring           Poisson(lambda) per infectant                           src/simulation.rs:148-152
//   k_plan_resample       target_size + sample size + split of the draws          src/simulation.rs:491-527
//   k_resample            Population::sample / sample_indices / select            src/core/population.rs:385-430
//   k_gc_*                Rc/Arc drop of unreachable haplotypes                   src/references/sync.rs:107-115
//
// Data flow of one generation: the counting sort by host does not only produce indices (the reference's
// `hosts` array) but moves a record (infectant index, haplotype id) to its CSR position.  From there on every
// pass streams over host-contiguous records: per-host sums are segment sums over neighbours, lambda and the
// offspring counts live at CSR positions, and the resample draws positions.  Random accesses per infectant are
// the histogram atomic, the offsets gather + record scatter, and the gather of the haplotype's raw fitness.
// All kernels are grid-stride over device-resident sizes (DevState::n, n_sorted, ...), so a whole generation is
// enqueued without the host ever reading a size back.
#pragma once
#include "vl_rng.cuh"
#include "vl_scan.cuh"
#include "vl_state.cuh"

namespace vl {

constexpr int TPB = 256;
constexpr uint32_t MAX_BUCKETS_C = 2048;  // = MAX_BUCKETS (declared with the bucketed kernels)
constexpr uint32_t THREAD_TIER = 32;      // slices up to this length are handled by one thread
constexpr uint32_t SMEM_SORT_MAX = 4096;  // slices up to this length are sorted in shared memory (u64 keys)
constexpr int INFECT_ITEMS = 4;           // infectants per thread and round in k_infect / k_scatter

__device__ __forceinline__ void raise(DevState *st, uint32_t bit) { atomicOr(&st->error, bit); }
// ---- L2 residency hints.  The per-host sums of the device-resident generation (H x 8 bytes, 80 MB at H = 10^7) are the
// only array a generation touches at random: they are accessed with an evict-last policy, the per-infectant streams that
// flow past them with evict-first, so that the sums stay in L2 instead of costing a DRAM sector per access.
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_stream_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// f64 reduction into global memory without a return value (RED, not ATOM: nothing waits for the old value)
__device__ __forceinline__ void red_add_f64(double *p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add_f64_hint(double *p, double v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(__cvta_generic_to_global(p)), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double ld_f64_hint(const double *p, uint64_t pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(__cvta_generic_to_global(p)), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_u32_hint(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(__cvta_generic_to_global(p)), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_f64_stream(const double *p, uint64_t pol) {
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(__cvta_generic_to_global(p)), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_u32_hint(uint32_t *p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(__cvta_generic_to_global(p)), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_f64_hint(double *p, double v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(__cvta_generic_to_global(p)), "d"(v), "l"(pol) : "memory");
}
// memoise a record's raw fitness under provider p together with the mapped value utility(raw): the raw value is what
// descendants multiply (FitnessProvider::get_fitness, src/providers/fitness.rs:212-223), the mapped value is what
// infect and replicate read (AttributeSet::get_or_compute, src/core/attributes.rs:220-241) — N reads per generation
__device__ __forceinline__ void set_fitness(DevState *st, uint32_t p, uint64_t id, double raw) {
    st->h_raw[p][id] = raw;
    st->h_fit[p][id] = utility_apply(st->prov[p], raw);
}
// head of a record that owns a flat list (no delta): recombinants, imported migrants, compacted records.  The fitness
// of every provider has been memoised before (set_fitness).
__device__ __forceinline__ void write_flat_head(DevState *st, uint64_t id, uint64_t off, uint32_t len) {
    HapHead h;
    h.off = off;
    h.len = len;
    h.nd_mlen = len << 8;
    h.raw0 = st->n_providers ? st->h_raw[0][id] : 1.0;
    h.fit0 = st->n_providers ? st->h_fit[0][id] : 1.0;
#pragma unroll
    for (uint32_t k = 0; k < HEAD_DELTA; k++) h.d[k] = 0xFFFFFFFFu;
    st->head[id] = h;
}
__global__ void k_init_wildtype(DevState *st) {  // haplotype 0: no entries, raw fitness = empty product (src/core/fitness/table.rs:72-79)
    for (uint32_t p = 0; p < st->n_providers; p++) set_fitness(st, p, 0, 1.0);
    write_flat_head(st, 0, 0, 0);
}
// A pointer read from DevState is a generic pointer to the compiler; every array it names lives in global memory.
// Saying so where the pointer is used turns generic LD / ST into LDG / STG without keeping the pointer in a register.
template <typename T>
__device__ __forceinline__ T *gptr(T *p) {
    __builtin_assume(__isGlobal(p));
    return p;
}
// the merged (flattened) record of `id` written to out[]; returns its length (= head_mlen)
__device__ inline uint32_t flatten_record(const DevState *st, const HapHead &h, uint32_t *out) {
    uint32_t o = 0;
    merged_for_each(st->arena + h.off, h.len, h.d, head_nd(h), [&](uint32_t e) { out[o++] = e; });
    return o;
}
__device__ __forceinline__ uint64_t vl_globaltimer() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// HostSpecs::try_get_spec_from_index (src/core/hosts.rs:62-64): first spec whose range contains the host
__device__ __forceinline__ int find_spec(const DevState *st, uint64_t h) {
    for (uint32_t s = 0; s < st->n_specs; s++)
        if (h >= st->specs[s].lo && h < st->specs[s].hi) return (int)s;
    return -1;
}

// exclusive scan of one u32 per thread over the CTA (TPB threads); returns the thread's offset, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t x, uint32_t *total) {
    __shared__ uint32_t s_w[TPB / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += y;
    }
    __syncthreads();  // protect s_w reuse across calls
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < TPB / 32; w++) {
        uint32_t t = s_w[w];
        if (w < (int)wid) woff += t;
        tot += t;
    }
    *total = tot;
    return woff + inc - x;
}

// ------------------------------------------------------------------------------------------ generation
__global__ void k_tick(DevState *st) {  // Simulation::increment_generation (src/simulation.rs:178)
    st->generation += 1;
}
__global__ void k_begin_phase(DevState *st) {
    st->n_mut_list = 0;
    st->n_heavy = 0;
    st->n_medium = 0;
}
__global__ void k_plan_host_tiles(DevState *st) { st->ht_planned = host_tile_size(st->n, st->H); }
// everything the fused generation clears or bumps before its first real kernel, in one launch
__global__ void __launch_bounds__(1024) k_begin_generation(DevState *st) {
    for (uint32_t b = threadIdx.x; b <= MAX_BUCKETS_C; b += blockDim.x) st->bucket_fill[b] = 0u;
    if (threadIdx.x == 0) {
        st->generation += 1;  // Simulation::increment_generation (src/simulation.rs:178)
        st->n_mut_list = 0;
        st->n_heavy = 0;
        st->n_medium = 0;
        st->fused_failed = 0;
        st->ht_planned = host_tile_size(st->n, st->H);
    }
}

// ------------------------------------------------------------------------------------------ infect
// BasicHost::infect with n_hits > 1, several specs or an infective provider (src/simulation.rs:56-65):
// sequential draws from block 1 onwards of the infectant's PH_INFECT stream.
__device__ __noinline__ uint32_t infect_general(const DevState *st, uint64_t i, uint32_t gen) {
    Philox g(st->seed, st->compartment, gen, PH_INFECT, i, 1);
    const uint64_t H = st->H;
    for (uint64_t hit = 0; hit < st->n_hits; hit++) {
        uint64_t cand = g.below(H);
        int s = find_spec(st, cand);
        if (s < 0) continue;
        double infectivity = 1.;
        int pi = st->specs[s].infect;
        if (pi >= 0) infectivity = st->h_fit[pi][st->hap_id[i]];
        if (g.u01() < infectivity) return (uint32_t)cand;
    }
    return NONE32;
}

// Mutation-count uniform of infectant i: words z,w of block 0 of its PH_INFECT stream (word x is the host draw
// of the single-spec fast path).  The same definition is used by the fused and the stand-alone count kernel.
__device__ __noinline__ uint32_t mutation_count_general(const DevState *st, uint64_t i, uint32_t gen) {
    Philox g(st->seed, st->compartment, gen, PH_MUTCOUNT, i);  // large-mean regime: BTRS
    return (uint32_t)binomial(g, st->L, st->host_mutation_rate);
}
__device__ __noinline__ uint32_t mutation_count_binv(const DevState *st, PhiloxKey key, uint64_t i, uint32_t gen, double u) {
    return binv_draw(st->binv, u, key, i, gen, PH_MUTCOUNT);
}
// small-mean regime only (BinvConst::ok): what the fused infect kernel calls
__device__ __forceinline__ uint32_t mutation_count_small(const DevState *st, PhiloxKey key, uint64_t i, uint32_t gen, uint4 b0) {
    const uint64_t U = (((uint64_t)b0.w << 32) | b0.z) >> 11;  // u01_from(b0.z, b0.w) = U * 2^-53
    if (U < st->binv.thr[0]) return 0;  // the common case: no mutation
    uint32_t k = 1;
    while (k < (uint32_t)BINV_THRESHOLDS && U >= st->binv.thr[k]) k++;
    if (k < (uint32_t)BINV_THRESHOLDS) return k;
    return mutation_count_binv(st, key, i, gen, u01_from(b0.z, b0.w));
}
__device__ __forceinline__ uint32_t mutation_count(const DevState *st, PhiloxKey key, uint64_t i, uint32_t gen, uint4 b0) {
    if (st->binv.ok) return mutation_count_small(st, key, i, gen, b0);
    return mutation_count_general(st, i, gen);
}

// append (infectant, n_mut) for every item with n_mut > 0: one global atomic per CTA and round
template <int ITEMS>
__device__ __forceinline__ void worklist_append(DevState *st, const uint64_t (&idx)[ITEMS], const uint32_t (&nm)[ITEMS], const uint32_t (&pos)[ITEMS]) {
    __shared__ uint32_t s_base;
    uint32_t mine = 0;
#pragma unroll
    for (int e = 0; e < ITEMS; e++) mine += nm[e] ? 1u : 0u;
    uint32_t tot;
    uint32_t off = block_excl_scan(mine, &tot);
    if (threadIdx.x == 0 && tot) s_base = atomicAdd(&st->n_mut_list, tot);
    __syncthreads();
    if (mine) {
        uint32_t o = s_base + off;
#pragma unroll
        for (int e = 0; e < ITEMS; e++)
            if (nm[e]) st->mut_list[o++] = make_uint4((uint32_t)idx[e], nm[e], pos[e], 0u);
    }
    __syncthreads();
}

// replay_inf: recorded Option<usize> per infectant (-1 = None) or nullptr.
// FAST: one spec over [0,H), no infective provider, n_hits == 1 (DevState::fast_infect).  FUSE_MUT needs BinvConst::ok.
template <bool FUSE_MUT, bool FAST>
__global__ void __launch_bounds__(TPB) k_infect(DevState *st, uint32_t *__restrict__ count, const int64_t *__restrict__ replay_inf) {
    const uint64_t n = st->n;
    const uint64_t H = st->H;
    const uint32_t H32 = (uint32_t)H;
    const uint32_t gen = st->generation;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    uint32_t *__restrict__ host_id = st->host_id;
    uint32_t *__restrict__ rank = st->rank;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->infectant_generations += n;
    const uint64_t span = (uint64_t)gridDim.x * blockDim.x * INFECT_ITEMS;
    const uint64_t rounds = (n + span - 1) / span;
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t i0 = r * span + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * INFECT_ITEMS;
        uint32_t h[INFECT_ITEMS], nm[INFECT_ITEMS], rk[INFECT_ITEMS];
        uint64_t idx[INFECT_ITEMS];
#pragma unroll
        for (int e = 0; e < INFECT_ITEMS; e++) {
            const uint64_t i = i0 + e;
            idx[e] = i;
            h[e] = NONE32;
            nm[e] = 0;
            rk[e] = 0;
            if (i >= n) continue;
            const uint4 b0 = philox_block(key, i, gen, PH_INFECT, 0);
            if (replay_inf) {
                int64_t v = replay_inf[i];
                if (v >= 0) {
                    if ((uint64_t)v < H) h[e] = (uint32_t)v; else raise(st, DE_REPLAY);
                }
            } else if (H > 0) {
                if (FAST) {
                    // Uniform(0, H): widening multiply; a draw in the biased zone (probability < H / 2^32) is
                    // rejected exactly and redrawn from the following blocks of the stream
                    uint64_t m = (uint64_t)b0.x * H32;
                    if ((uint32_t)m < H32) {
                        const uint32_t thr = (uint32_t)(-H32) % H32;
                        for (uint32_t blk = 1; (uint32_t)m < thr; blk++) m = (uint64_t)philox_block(key, i, gen, PH_INFECT, blk).x * H32;
                    }
                    h[e] = (uint32_t)(m >> 32);
                } else {
                    h[e] = infect_general(st, i, gen);
                }
            }
            if (FUSE_MUT && h[e] != NONE32) nm[e] = mutation_count_small(st, key, i, gen, b0);
        }
#pragma unroll
        for (int e = 0; e < INFECT_ITEMS; e++)
            if (h[e] != NONE32) rk[e] = atomicAdd(&count[h[e]], 1u);
        if (i0 + INFECT_ITEMS <= n) {
            *reinterpret_cast<uint4 *>(host_id + i0) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4 *>(rank + i0) = make_uint4(rk[0], rk[1], rk[2], rk[3]);
        } else {
#pragma unroll
            for (int e = 0; e < INFECT_ITEMS; e++)
                if (i0 + e < n) { host_id[i0 + e] = h[e]; rank[i0 + e] = rk[e]; }
        }
        if (FUSE_MUT) {
            const uint32_t none[INFECT_ITEMS] = {NONE32, NONE32, NONE32, NONE32};
            worklist_append<INFECT_ITEMS>(st, idx, nm, none);
        }
    }
}

// ---- bucketed variant (uniform host draws only: DevState::fast_infect) --------------------------------------
// Instead of a histogram atomic per infectant on H counters spread over HBM, the record (index, haplotype, host)
// is appended to the region of its host bucket (BUCKET_HOSTS consecutive hosts).  A CTA handles a chunk of
// BUCKET_CHUNK infectants per round: ranks inside the chunk come from a shared-memory histogram over the buckets,
// one global atomic per (chunk, bucket) reserves the run.  Every bucket region is an append-only stream, so the
// scattered 4-byte stores merge in L2 and reach HBM as full sectors.
constexpr int BUCKET_TPB = 1024;
constexpr int BUCKET_ITEMS = 8;
constexpr int BUCKET_CHUNK = BUCKET_TPB * BUCKET_ITEMS;
constexpr int MAX_BUCKETS = 2048;

// STORE_HOST: keep host_id[] (the reference's `infections`) for the trait-level phases and inspection calls; the fused
// generation does not read it back.
template <bool FUSE_MUT, bool STORE_HOST, int BTPB>
__global__ void __launch_bounds__(BTPB) k_infect_bucket(DevState *st) {
    __shared__ uint32_t s_hist[MAX_BUCKETS];   // per-chunk count, then global base of the chunk's run per bucket
    __shared__ uint32_t s_hap[BTPB * BUCKET_ITEMS];  // the chunk's haplotype ids, copied asynchronously while the hosts are drawn
    __shared__ uint32_t s_w[BTPB / 32];
    __shared__ uint32_t s_lbase;
    const uint64_t n = st->n;
    const uint32_t H32 = (uint32_t)st->H;
    const uint32_t gen = st->generation;
    const uint32_t B = st->n_buckets;
    const uint64_t stride = bucket_stride(n, st->H);
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    uint32_t *__restrict__ host_id = st->host_id;
    const uint32_t *__restrict__ hap_id = st->hap_id;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->infectant_generations += n;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t n_chunks = (n + (BTPB * BUCKET_ITEMS) - 1) / (BTPB * BUCKET_ITEMS);
    for (uint64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        for (uint32_t b = tid; b < B; b += BTPB) s_hist[b] = 0;
        // every thread starts the copy of its own items' haplotype ids into shared memory (cp.async: no register is held,
        // the latency hides under the Philox rounds below) and is the only reader of those words later
#pragma unroll
        for (int e = 0; e < BUCKET_ITEMS; e++) {
            const uint64_t i = chunk * (BTPB * BUCKET_ITEMS) + (uint64_t)e * BTPB + tid;
            if (i < n) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_hap[e * BTPB + tid]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(hap_id + i) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        __syncthreads();
        // coalesced layout: item e of thread t is infectant chunk*CHUNK + e*TPB + t.  Per item two registers survive
        // the barriers: the host, and (rank inside the chunk's bucket run | n_mut << 16).
        uint32_t h[BUCKET_ITEMS], rn[BUCKET_ITEMS];
#pragma unroll
        for (int e = 0; e < BUCKET_ITEMS; e++) {
            const uint64_t i = chunk * (BTPB * BUCKET_ITEMS) + (uint64_t)e * BTPB + tid;
            h[e] = NONE32;
            rn[e] = 0;
            if (i >= n) continue;
            const uint4 b0 = philox_block(key, i, gen, PH_INFECT, 0);
            uint64_t m = (uint64_t)b0.x * H32;
            if ((uint32_t)m < H32) {
                const uint32_t thr = (uint32_t)(-H32) % H32;
                for (uint32_t blk = 1; (uint32_t)m < thr; blk++) m = (uint64_t)philox_block(key, i, gen, PH_INFECT, blk).x * H32;
            }
            h[e] = (uint32_t)(m >> 32);
            if (STORE_HOST) host_id[i] = h[e];
            uint32_t nm = 0;
            if (FUSE_MUT) nm = min(mutation_count_small(st, key, i, gen, b0), 0xFFFFu);
            rn[e] = atomicAdd(&s_hist[h[e] >> BUCKET_SHIFT], 1u) | (nm << 16);
        }
        __syncthreads();
        // one global atomic per (chunk, bucket): the run of this chunk inside the bucket's region
        for (uint32_t b = tid; b < B; b += BTPB) {
            const uint32_t c = s_hist[b];
            s_hist[b] = c ? atomicAdd(&st->bucket_fill[b], c) : 0u;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        uint32_t mine = 0;
#pragma unroll
        for (int e = 0; e < BUCKET_ITEMS; e++) {
            if (h[e] == NONE32) continue;
            const uint64_t i = chunk * (BTPB * BUCKET_ITEMS) + (uint64_t)e * BTPB + tid;
            const uint32_t b = h[e] >> BUCKET_SHIFT;
            const uint64_t r = (uint64_t)s_hist[b] + (rn[e] & 0xFFFFu);
            const uint32_t nm = rn[e] >> 16;
            uint32_t pos = NONE32;
            if (r < stride) {
                pos = (uint32_t)(b * stride + r);
                st->tmp_rec[pos] = make_uint4((uint32_t)i, s_hap[e * BTPB + tid], h[e], 0u);
            } else {
                raise(st, DE_BUCKET_OVERFLOW);
            }
            h[e] = pos;   // the host is not needed any more: keep the record position for the worklist
            rn[e] = nm;
            mine += nm ? 1u : 0u;
        }
        if (FUSE_MUT) {
            // mutation worklist: (infectant, n_mut, position of its partitioned record); one global atomic per chunk
            uint32_t inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= (uint32_t)d) inc += y;
            }
            if (lane == 31) s_w[wid] = inc;
            __syncthreads();
            uint32_t woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < BTPB / 32; w++) {
                const uint32_t t = s_w[w];
                if (w < (int)wid) woff += t;
                tot += t;
            }
            if (tid == 0) s_lbase = tot ? atomicAdd(&st->n_mut_list, tot) : 0u;
            __syncthreads();
            if (mine) {
                uint32_t o = s_lbase + woff + inc - mine;
#pragma unroll
                for (int e = 0; e < BUCKET_ITEMS; e++)
                    if (rn[e]) st->mut_list[o++] = make_uint4((uint32_t)(chunk * (BTPB * BUCKET_ITEMS) + (uint64_t)e * BTPB + tid), rn[e], h[e], 0u);
            }
        }
        __syncthreads();
    }
}

// first CSR position of every bucket (exclusive scan of the fills), n_sorted = total.  One CTA, two buckets per thread.
__global__ void __launch_bounds__(MAX_BUCKETS / 2) k_bucket_offsets(DevState *st) {
    __shared__ uint32_t s_w[MAX_BUCKETS / 64];
    const uint32_t B = st->n_buckets, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t x0 = 2 * tid < B ? st->bucket_fill[2 * tid] : 0u, x1 = 2 * tid + 1 < B ? st->bucket_fill[2 * tid + 1] : 0u;
    const uint32_t x = x0 + x1;
    uint32_t inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += y;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
    for (int w = 0; w < MAX_BUCKETS / 64; w++) {
        const uint32_t t = s_w[w];
        if (w < (int)wid) woff += t;
        tot += t;
    }
    const uint32_t excl = woff + inc - x;
    if (2 * tid < B) st->bucket_pos[2 * tid] = excl;
    if (2 * tid + 1 < B) st->bucket_pos[2 * tid + 1] = excl + x0;
    if (tid == 0) { st->bucket_pos[B] = tot; st->n_sorted = tot; st->offsets[st->H] = tot; }
}

// One CTA per bucket: counting sort of the bucket's records by host with the histogram in shared memory.
// The records are staged in shared memory as well when the bucket fits (one global read per record); writes
// offsets[] for the bucket's hosts and the records at their CSR positions (a window of ~100 KB that stays in L2
// until its sectors are complete).
// SRECS > 0: records of a bucket that fits are staged in shared memory between the two passes (one CTA per SM);
// SRECS == 0: the second pass re-reads them through L2, which leaves room for several CTAs per SM.
constexpr uint32_t BSORT_STAGE_RECS = 11264;  // 11264 * 10 B = 110 KB next to the 32 KB histogram
template <int STPB, uint32_t SRECS>
__global__ void __launch_bounds__(STPB) k_bucket_sort(DevState *st) {
    extern __shared__ uint32_t s_dyn32[];
    uint32_t *s_cnt = s_dyn32;                                               // BUCKET_HOSTS counters, then running cursors
    uint2 *s_rec = reinterpret_cast<uint2 *>(s_dyn32 + BUCKET_HOSTS);        // SRECS (index, haplotype)
    uint16_t *s_hl = reinterpret_cast<uint16_t *>(s_rec + SRECS);  // SRECS host - first host of the bucket
    __shared__ uint32_t s_w[STPB / 32];
    const uint32_t B = st->n_buckets;
    const uint64_t H = st->H;
    const uint64_t stride = bucket_stride(st->n, H);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr uint32_t SEG = BUCKET_HOSTS / (STPB / 32);  // hosts scanned per warp
    for (uint32_t b = blockIdx.x; b < B; b += gridDim.x) {
        const uint32_t fill = min(st->bucket_fill[b], (uint32_t)stride);
        const uint4 *__restrict__ src = st->tmp_rec + (uint64_t)b * stride;
        const uint32_t h0 = b << BUCKET_SHIFT;
        const uint32_t pos0 = st->bucket_pos[b];
        const uint32_t nh = (uint32_t)min((uint64_t)BUCKET_HOSTS, H - h0);
        const bool staged = fill <= SRECS;
        for (uint32_t k = tid; k < BUCKET_HOSTS; k += STPB) s_cnt[k] = 0;
        __syncthreads();
        for (uint32_t j0 = tid; j0 < fill; j0 += 8 * STPB) {  // eight independent loads in flight per thread
            uint4 r[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t j = j0 + u * STPB;
                r[u] = j < fill ? src[j] : make_uint4(0u, 0u, h0, 0u);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t j = j0 + u * STPB;
                if (j >= fill) continue;
                const uint32_t hl = r[u].z - h0;
                atomicAdd(&s_cnt[hl], 1u);
                if (staged) { s_rec[j] = make_uint2(r[u].x, r[u].y); s_hl[j] = (uint16_t)hl; }
            }
        }
        __syncthreads();
        // exclusive scan: each warp scans a contiguous segment 32 hosts at a time (conflict-free), then segment bases
        uint32_t carry = 0;
        for (uint32_t it = 0; it < SEG / 32; it++) {
            const uint32_t k = wid * SEG + it * 32 + lane;
            const uint32_t x = s_cnt[k];
            uint32_t inc = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= (uint32_t)d) inc += y;
            }
            s_cnt[k] = carry + inc - x;
            carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (lane == 0) s_w[wid] = carry;
        __syncthreads();
        uint32_t wbase = 0;
        for (uint32_t w = 0; w < wid; w++) wbase += s_w[w];
        for (uint32_t it = 0; it < SEG / 32; it++) {
            const uint32_t k = wid * SEG + it * 32 + lane;
            const uint32_t o = s_cnt[k] + wbase;
            s_cnt[k] = o;
            if (k < nh) st->offsets[h0 + k] = pos0 + o;
        }
        __syncthreads();
        if (staged) {
            for (uint32_t j = tid; j < fill; j += STPB) st->rec[pos0 + atomicAdd(&s_cnt[s_hl[j]], 1u)] = s_rec[j];
        } else {
            for (uint32_t j0 = tid; j0 < fill; j0 += 4 * STPB) {  // second read: L2 hits, four in flight per thread
                uint4 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t j = j0 + u * STPB;
                    q[u] = j < fill ? src[j] : make_uint4(0u, 0u, h0, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t j = j0 + u * STPB;
                    if (j < fill) st->rec[pos0 + atomicAdd(&s_cnt[q[u].z - h0], 1u)] = make_uint2(q[u].x, q[u].y);
                }
            }
        }
        __syncthreads();
    }
}

// HostMapBuffer::build back-fill (src/core/hosts.rs:182-189): the record of infectant i goes to its host's slice
__global__ void __launch_bounds__(TPB) k_scatter(DevState *st) {
    const uint64_t n = st->n;
    const uint32_t *__restrict__ host_id = st->host_id;
    const uint32_t *__restrict__ rank = st->rank;
    const uint32_t *__restrict__ hap_id = st->hap_id;
    const uint32_t *__restrict__ offsets = st->offsets;
    uint2 *__restrict__ rec = st->rec;
    const uint64_t span = (uint64_t)gridDim.x * blockDim.x * INFECT_ITEMS;
    for (uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * INFECT_ITEMS; i0 < n; i0 += span) {
        uint32_t h[INFECT_ITEMS], rk[INFECT_ITEMS], hp[INFECT_ITEMS], ob[INFECT_ITEMS];
        if (i0 + INFECT_ITEMS <= n) {
            uint4 a = *reinterpret_cast<const uint4 *>(host_id + i0), b = *reinterpret_cast<const uint4 *>(rank + i0),
                  c = *reinterpret_cast<const uint4 *>(hap_id + i0);
            h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w;
            rk[0] = b.x; rk[1] = b.y; rk[2] = b.z; rk[3] = b.w;
            hp[0] = c.x; hp[1] = c.y; hp[2] = c.z; hp[3] = c.w;
        } else {
#pragma unroll
            for (int e = 0; e < INFECT_ITEMS; e++) {
                const bool ok = i0 + e < n;
                h[e] = ok ? host_id[i0 + e] : NONE32;
                rk[e] = ok ? rank[i0 + e] : 0;
                hp[e] = ok ? hap_id[i0 + e] : 0;
            }
        }
#pragma unroll
        for (int e = 0; e < INFECT_ITEMS; e++) ob[e] = h[e] != NONE32 ? offsets[h[e]] : 0;
#pragma unroll
        for (int e = 0; e < INFECT_ITEMS; e++)
            if (h[e] != NONE32) rec[ob[e] + rk[e]] = make_uint2((uint32_t)(i0 + e), hp[e]);
    }
}

// identity layout (no host map): record p = infectant p.  Used when an offspring map comes from the host.
__global__ void __launch_bounds__(TPB) k_identity_records(DevState *st) {
    const uint64_t n = st->n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) st->rec[i] = make_uint2((uint32_t)i, st->hap_id[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->n_sorted = n; st->host_tiles = 0; }
}

// offspring counts moved from CSR positions to infectant order (tmp[i], zero for uninfected infectants): every path
// resamples over tiles of RESAMPLE_TILE consecutive infectants, so the draws of a generation do not depend on which
// kernels built its offspring map
__global__ void __launch_bounds__(TPB) k_identity_from(DevState *st, const uint32_t *__restrict__ tmp) {
    const uint64_t n = st->n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        st->offspring[i] = tmp[i];
        st->rec[i] = make_uint2((uint32_t)i, st->hap_id[i]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->n_sorted = n; st->host_tiles = 0; }
}

// ------------------------------------------------------------------------------------------ haplotype records
struct NewRecord {
    uint32_t id;
    uint64_t off;
    bool ok;
};

// block-aggregated reservation of `words` arena words and one haplotype id for every thread with want=true
__device__ __forceinline__ NewRecord reserve_record(DevState *st, bool want, uint32_t words) {
    __shared__ unsigned long long s_abase;
    __shared__ unsigned long long s_ibase;
    uint32_t wtot, itot;
    uint32_t woff = block_excl_scan(want ? words : 0u, &wtot);
    uint32_t ioff = block_excl_scan(want ? 1u : 0u, &itot);
    if (threadIdx.x == 0) {
        s_abase = itot ? atomicAdd(&st->arena_top, (unsigned long long)wtot) : 0ull;
        s_ibase = itot ? atomicAdd(&st->n_hap, (unsigned long long)itot) : 0ull;
    }
    __syncthreads();
    NewRecord r;
    r.off = s_abase + woff;
    const uint64_t id64 = s_ibase + ioff;
    r.id = (uint32_t)id64;
    r.ok = want;
    if (want) {
        if (id64 >= st->cap_hap) { raise(st, DE_HAP_CAPACITY); r.ok = false; }
        if (r.off + words > st->cap_arena) { raise(st, DE_ARENA_CAPACITY); r.ok = false; }
    }
    __syncthreads();
    return r;
}
// single-thread variant (rare events: recombination)
__device__ __forceinline__ NewRecord reserve_record_single(DevState *st, uint32_t words) {
    NewRecord r;
    r.off = atomicAdd(&st->arena_top, (unsigned long long)words);
    const uint64_t id64 = atomicAdd(&st->n_hap, 1ull);
    r.id = (uint32_t)id64;
    r.ok = true;
    if (id64 >= st->cap_hap) { raise(st, DE_HAP_CAPACITY); r.ok = false; }
    if (r.off + words > st->cap_arena) { raise(st, DE_ARENA_CAPACITY); r.ok = false; }
    return r;
}

__device__ inline void heap_sort_u32(uint32_t *a, uint32_t n) {
    for (uint32_t start = n / 2; start-- > 0;) {
        uint32_t root = start;
        for (;;) {
            uint32_t child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            uint32_t t = a[root]; a[root] = a[child]; a[child] = t;
            root = child;
        }
    }
    for (uint32_t end = n; end-- > 1;) {
        uint32_t t = a[0]; a[0] = a[end]; a[end] = t;
        uint32_t root = 0;
        for (;;) {
            uint32_t child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            uint32_t t2 = a[root]; a[root] = a[child]; a[child] = t2;
            root = child;
        }
    }
}

// change word in scratch: pos << 8 | from << 4 | to
__device__ __forceinline__ uint32_t ch_pos(uint32_t c) { return c >> 8; }
__device__ __forceinline__ uint32_t ch_from(uint32_t c) { return (c >> 4) & 0x7; }
__device__ __forceinline__ uint32_t ch_present(uint32_t c) { return (c >> 7) & 1u; }  // the pre-mutation haplotype had an entry at this site
__device__ __forceinline__ uint32_t ch_to(uint32_t c) { return c & 0xF; }

// EpistasisTable::update_fitness (src/core/fitness/epistasis.rs:105-173) on a freshly merged record.
// Quirks kept: candidates are the changes whose (pos,to) has a row (`from` is never tested); entries of the
// mutation set at positions <= the change that belong to this node's own changes are skipped.
// The reference walks the haplotype's whole mutation set per change (O(length)).  The factors that apply are the entries of
// the change's two rows whose partner (position, base) is in the set, so when the rows are the shorter side (a site has a
// handful of partners, a haplotype hundreds of mutations after a long run) the rows are walked instead, both in ascending
// position, each candidate looked up in the set: the same factors in the same order, bit for bit, at a cost that does
// not grow with the haplotype.
__device__ inline double epi_update(const DevProvider &p, const uint32_t *chs, uint32_t n_ch, const uint32_t *base, uint32_t len, const uint32_t *d, uint32_t nd,
                                    uint32_t L) {
    bool any = false;
    for (uint32_t k = 0; k < n_ch && !any; k++) {
        uint32_t lo, hi;
        epi_row(p, ekey(ch_pos(chs[k]), ch_to(chs[k])), lo, hi);
        any = hi > lo;
    }
    if (!any) return 1.;
    double fitness = 1.;
    for (uint32_t k = 0; k < n_ch; k++) {
        const uint32_t cpos = ch_pos(chs[k]);
        uint32_t alo, ahi, rlo, rhi;
        epi_row(p, ekey(cpos, ch_to(chs[k])), alo, ahi);
        if (ahi == alo) continue;
        epi_row(p, ekey(cpos, ch_from(chs[k])), rlo, rhi);
        if ((ahi - alo) + (rhi - rlo) <= len + nd) {
            uint32_t a = alo, r = rlo;
            while (a < ahi || r < rhi) {
                const uint32_t ka = a < ahi ? p.epi[a].k2 : 0xFFFFFFFFu, kr = r < rhi ? p.epi[r].k2 : 0xFFFFFFFFu;
                const uint32_t pos = (ka < kr ? ka : kr) >> 2;
                const uint32_t e = merged_find(base, len, d, nd, pos, L);
                bool live = e != 0xFFFFFFFFu && entry_m(e) < 4u && pos != cpos;
                if (live && pos <= cpos)
                    for (uint32_t q = 0; q < n_ch; q++) if (ch_pos(chs[q]) == pos) live = false;  // one of this node's own changes
                const uint32_t key = live ? ekey(pos, entry_m(e)) : 0xFFFFFFFFu;
                for (; a < ahi && (p.epi[a].k2 >> 2) == pos; a++) if (p.epi[a].k2 == key) fitness *= p.epi[a].v;
                for (; r < rhi && (p.epi[r].k2 >> 2) == pos; r++) if (p.epi[r].k2 == key) fitness /= p.epi[r].v;
            }
            continue;
        }
        // the new record's mutation set M in ascending position (the reference walks a HashMap; the factors commute up to rounding)
        merged_for_each(base, len, d, nd, [&](uint32_t e) {
            const uint32_t mb = entry_m(e);
            if (mb >= 4) return;
            const uint32_t pos = entry_pos(e);
            if (pos <= cpos) {
                for (uint32_t q = 0; q < n_ch; q++) if (ch_pos(chs[q]) == pos) return;  // one of this node's own changes
            }
            if (pos != cpos) {
                double v;
                if (epi_get(p, alo, ahi, ekey(pos, mb), v)) fitness *= v;
                if (rhi > rlo && epi_get(p, rlo, rhi, ekey(pos, mb), v)) fitness /= v;
            }
        });
    }
    return fitness;
}

// ------------------------------------------------------------------------------------------ slice order
// The reference back-fills each host's slice while walking infectants upwards, which leaves the slice in
// DESCENDING infectant order (KAT src/core/hosts.rs:317-318).  Records arrive in the arrival order of the
// histogram atomics; sorting them makes the layout (and everything keyed by CSR position) deterministic.
// ------------------------------------------------------------------------------------------ recombination
struct ReplayRecomb {
    uint64_t n_events;
    const uint64_t *host;  // sorted ascending; events of one host in processing order
    const uint32_t *i, *j, *s0, *s1;
    const uint8_t *which;
};

// Haplotype::create_recombinant (src/core/haplotype.rs:269-288) flattened: left entries inside [s0,s1),
// right entries outside; raw fitness by full recomputation like FitnessProvider::get_fitness does for
// nodes without changes (src/providers/fitness.rs:224-227).
__device__ inline uint32_t make_recombinant(DevState *st, uint32_t left, uint32_t right, uint32_t s0, uint32_t s1) {
    const HapHead hl = st->head[left], hr = st->head[right];
    NewRecord r = reserve_record_single(st, head_mlen(hl) + head_mlen(hr));
    if (!r.ok) return left;
    uint32_t *out = st->arena + r.off;
    uint32_t o = 0;
    // right entries below s0, left entries inside [s0, s1), right entries from s1 on: three passes over the merged views
    merged_for_each(st->arena + hr.off, hr.len, hr.d, head_nd(hr), [&](uint32_t e) { if (entry_pos(e) < s0) out[o++] = e; });
    merged_for_each(st->arena + hl.off, hl.len, hl.d, head_nd(hl), [&](uint32_t e) { const uint32_t p = entry_pos(e); if (p >= s0 && p < s1) out[o++] = e; });
    merged_for_each(st->arena + hr.off, hr.len, hr.d, head_nd(hr), [&](uint32_t e) { if (entry_pos(e) >= s1) out[o++] = e; });
    for (uint32_t p = 0; p < st->n_providers; p++) set_fitness(st, p, r.id, compute_fitness_full(st->prov[p], out, o));
    write_flat_head(st, r.id, r.off, o);
    return r.id;
}

// One thread per host slice (recombination is sequential inside a host: later pairs see earlier replacements).
// Needs ordered slices: runs after k_host_order.
__global__ void __launch_bounds__(TPB) k_recombine(DevState *st, ReplayRecomb rp, int replay) {
    const uint64_t H = st->H;
    const uint32_t L = st->L;
    const double rho = st->host_recombination_rate;
    const uint32_t *__restrict__ offsets = st->offsets;
    uint2 *rec = st->rec;
    uint32_t *hap_id = st->hap_id;
    const double log1m = (rho < 1.0) ? log1p(-rho) : 0.0;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < H; h += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = offsets[h], k = offsets[h + 1] - b;
        if (k <= 1) continue;
        if (replay) {
            uint64_t lo = 0, hi = rp.n_events;
            while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (rp.host[mid] < h) lo = mid + 1; else hi = mid; }
            for (uint64_t e = lo; e < rp.n_events && rp.host[e] == h; e++) {
                uint32_t i = rp.i[e], j = rp.j[e];
                if (i >= k || j >= k || rp.s0[e] >= rp.s1[e] || rp.s1[e] > L) { raise(st, DE_REPLAY); continue; }
                uint32_t recid = make_recombinant(st, rec[b + i].y, rec[b + j].y, rp.s0[e], rp.s1[e]);
                const uint32_t w = b + (rp.which[e] ? j : i);
                rec[w].y = recid;
                hap_id[rec[w].x] = recid;
            }
            continue;
        }
        // pairs (i<j) in lexicographic order (itertools tuple_combinations); Bernoulli(rho) per pair is
        // drawn as geometric gaps between successes — same joint law, O(#events) instead of O(k^2)
        Philox g(st->seed, st->compartment, st->generation, PH_RECOMB, h);
        const uint64_t K = (uint64_t)k * (k - 1) / 2;
        uint64_t pos = 0;  // index of the next pair to be tested
        uint32_t row = 0;
        uint64_t row_start = 0;  // linear index of pair (row, row+1)
        for (;;) {
            uint64_t skip = 0;
            if (rho < 1.0) {
                double u = g.u01_open();
                double s = floor(log(u) / log1m);
                if (!(s < 1.8e19)) break;
                skip = (uint64_t)s;
            }
            if (skip >= K - pos) break;
            pos += skip;
            while (pos >= row_start + (k - 1 - row)) { row_start += (k - 1 - row); row++; }
            const uint32_t i = row, j = row + 1 + (uint32_t)(pos - row_start);
            uint32_t s0 = (uint32_t)g.below(L), s1 = (uint32_t)g.below(L);
            while (s0 == s1) s1 = (uint32_t)g.below(L);
            if (s0 > s1) { uint32_t t = s0; s0 = s1; s1 = t; }
            uint32_t recid = make_recombinant(st, rec[b + i].y, rec[b + j].y, s0, s1);
            const uint32_t w = b + ((uint32_t)g.below(2) ? j : i);
            rec[w].y = recid;
            hap_id[rec[w].x] = recid;
            pos += 1;
            if (pos >= K) break;
        }
    }
}

// ------------------------------------------------------------------------------------------ mutation
struct ReplayMut {
    const uint64_t *offsets;  // n+1
    const uint32_t *sites;
    const uint8_t *bases;
};

// Stand-alone mutation counts (replayed generations, or infections that were replayed without a mutation log):
// n_mut per infected infectant, the ones with n_mut > 0 go to the worklist.
__global__ void __launch_bounds__(TPB) k_mutation_counts(DevState *st, ReplayMut rp, int replay) {
    const uint64_t n = st->n;
    const uint32_t *__restrict__ host_id = st->host_id;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const uint32_t gen = st->generation;
    const uint64_t span = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + span - 1) / span;
    for (uint64_t r = 0; r < rounds; r++) {
        uint64_t idx[1] = {r * span + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x};
        uint32_t nm[1] = {0};
        const uint32_t pos[1] = {NONE32};
        const uint64_t i = idx[0];
        if (i < n && host_id[i] != NONE32) {
            if (replay) nm[0] = (uint32_t)(rp.offsets[i + 1] - rp.offsets[i]);
            else nm[0] = mutation_count(st, key, i, gen, philox_block(key, i, gen, PH_INFECT, 0));
        } else if (i < n && replay && rp.offsets[i + 1] != rp.offsets[i]) {
            raise(st, DE_REPLAY);  // the reference never mutates an uninfected infectant
        }
        worklist_append<1>(st, idx, nm, pos);
    }
}

// Mutation worklist of a uniformly infected population (DevState::fast_infect: every infectant is infected): the count
// of infectant i is a function of (seed, generation, i) alone — words z,w of block 0 of its PH_INFECT stream, the same
// uniform the fused infect kernels use — so the fused generation can mutate BEFORE it builds the host map and no
// record needs patching afterwards.  One global atomic per CTA and round.
constexpr int MC_ITEMS = 4;
__global__ void __launch_bounds__(TPB) k_mutcount_fast(DevState *st) {
    __shared__ uint32_t s_base;
    const uint64_t n = st->n;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const uint32_t gen = st->generation;
    const uint64_t chunk = (uint64_t)TPB * MC_ITEMS;
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    const uint32_t *__restrict__ hap_id = st->hap_id;
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        uint32_t nm[MC_ITEMS], hp[MC_ITEMS];
        uint32_t mine = 0;
#pragma unroll
        for (int e = 0; e < MC_ITEMS; e++) {  // the haplotype id travels with the item: read here it is a coalesced stream
            const uint64_t i = c * chunk + (uint64_t)e * TPB + threadIdx.x;
            hp[e] = i < n ? hap_id[i] : 0u;
        }
#pragma unroll
        for (int e = 0; e < MC_ITEMS; e++) {
            const uint64_t i = c * chunk + (uint64_t)e * TPB + threadIdx.x;
            nm[e] = 0;
            if (i < n) nm[e] = mutation_count_small(st, key, i, gen, philox_block(key, i, gen, PH_INFECT, 0));
            mine += nm[e] ? 1u : 0u;
        }
        uint32_t tot;
        const uint32_t off = block_excl_scan(mine, &tot);
        if (threadIdx.x == 0 && tot) s_base = atomicAdd(&st->n_mut_list, tot);
        __syncthreads();
        if (mine) {
            uint32_t o = s_base + off;
#pragma unroll
            for (int e = 0; e < MC_ITEMS; e++)
                if (nm[e]) st->mut_list[o++] = make_uint4((uint32_t)(c * chunk + (uint64_t)e * TPB + threadIdx.x), nm[e], NONE32, hp[e]);
        }
        __syncthreads();
    }
}

// One thread per mutated infectant: draw sites (with replacement) and sort them, look up the current base on the
// pre-mutation haplotype (its delta, then its base list), draw the substitution, and derive the child's record from the
// parent's 64-byte line: the changes are merged into the inline delta, or — when the delta would overflow — the record is
// flattened into a fresh base list (HapHead, vl_state.cuh).  Raw fitness of every provider incrementally:
// raw(parent) * fold(acc * T[pos,to] / T[pos,from]) (src/core/fitness/table.rs:65-68) [* epistatic update].
// patch_records: 1 = the infectant's CSR record is patched in place, 2 = its bucket-partitioned record is (host map not
// built yet), 0 = no records exist, 3 / 4 = device-resident generation (vl_fused.cuh): items are (infectant, n_mut, host,
// parent); the carried fitness of the infectant is replaced and added to its host's sum.
constexpr uint32_t MUT_LOCAL = HEAD_DELTA;              // events up to this many sites keep their scratch in shared memory
constexpr uint32_t MUT_STRIDE = 3 * HEAD_DELTA + 2 * MUT_LOCAL + 1;  // per thread: parent delta (8), merged delta (16), change words (8), new entries (8); odd stride
constexpr uint32_t MUT_WT_MAX_SITES = 65536;           // wildtypes up to this length are staged in shared memory, 2 bits per site
__host__ __device__ inline uint32_t mutate_wt_words(uint32_t L) { return L <= MUT_WT_MAX_SITES ? (L + 15u) / 16u : 0u; }
__host__ __device__ inline uint32_t mutate_smem_bytes(uint32_t L) { return (TPB * MUT_STRIDE + mutate_wt_words(L)) * 4u; }

// One flatten job done by the whole warp: out[] = base list (pe, plen) overridden by the merged delta dmv[0..ndm),
// ndm <= 32, tombstones dropped; returns the number of entries written.  Every lane j < ndm places its delta entry
// (lower bound in the base list by interpolation), the base entries are copied 32 at a time: nothing is serial in the
// list's length.
__device__ __forceinline__ uint32_t warp_flatten(const uint32_t *pe, uint32_t plen, const uint32_t *dmv, uint32_t ndm, uint32_t *out, uint32_t L) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t e = 0, bj = 0;
    bool hit = false, keep = false;
    if (lane < ndm) {  // every delta entry finds its place in the base list (one memory latency, all entries at once)
        e = dmv[lane];
        bj = entries_lower_interp(pe, plen, entry_pos(e), L);
        hit = bj < plen && entry_pos(pe[bj]) == entry_pos(e);
        keep = !entry_is_tombstone(e);
    }
    const uint32_t hitmask = __ballot_sync(0xFFFFFFFFu, hit), keepmask = __ballot_sync(0xFFFFFFFFu, keep);
    const uint32_t below = (1u << lane) - 1u;
    if (keep) out[bj - __popc(hitmask & below) + __popc(keepmask & below)] = e;
    // every base entry finds the delta entries before it (bisection over <= 32 positions in shared memory); four chunks of
    // 32 entries are in flight at a time
    for (uint32_t a0 = 0; a0 < plen; a0 += 128) {
        uint32_t be[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t a = a0 + u * 32 + lane;
            be[u] = a < plen ? pe[a] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t a = a0 + u * 32 + lane;
            if (a >= plen) continue;
            const uint32_t q = entry_pos(be[u]);
            uint32_t lo = 0, hi = ndm;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (entry_pos(dmv[mid]) < q) lo = mid + 1; else hi = mid;
            }
            if (lo < ndm && entry_pos(dmv[lo]) == q) continue;  // overridden (or removed) by the delta
            const uint32_t m = (1u << lo) - 1u;
            out[a - __popc(hitmask & m) + __popc(keepmask & m)] = be[u];
        }
    }
    return plen - __popc(hitmask) + __popc(keepmask);
}

// parent_in_item: the worklist items carry the infectant's current haplotype id in .w
// range: 0 = the whole worklist.  1 = items [0, mut_own) AHEAD of their generation: the population they belong to is still
// being assembled (hap_id_next / pop_fit_next / the other hsum buffer, generation + 1), their ids were reserved
// (mut_own_base), other kernels create records meanwhile.  2 = the items behind those, once the generation has begun.
// EPI: some provider is epistatic (the pair-table code is compiled into this instantiation only: the others stay leaner).
// FUSED: launched by the device-resident generation (no replay log, items carry parent and host, patch_records 3 or 4):
// the replay and host-map patching code is compiled out.
// BIGMODE: 0 = events of any size are handled where they are met (scratch of the large ones in the arena); 1 = events of
// more than MUT_LOCAL sites are set aside (DevState::big_list) and every scratch pointer is a shared-memory pointer;
// 2 = the pass that works off the list (same ids: a record's id is its worklist position).
template <bool EPI, bool FUSED = false, int BIGMODE = 0>
__global__ void __launch_bounds__(TPB, 4) k_mutate(DevState *st, ReplayMut rp, int replay_arg, int patch_records_arg, int parent_in_item_arg, int range) {
    const int replay = FUSED ? 0 : replay_arg;
    const int patch_records = FUSED ? (patch_records_arg == 4 ? 4 : 3) : patch_records_arg;
    const int parent_in_item = FUSED ? 1 : parent_in_item_arg;
    if (BIGMODE == 2 && st->n_big == 0u) return;  // (nearly always: nothing was set aside)
    extern __shared__ uint32_t s_mut_dyn[];
    __shared__ double s_subst[16];
    uint32_t *s_scratch = s_mut_dyn;
    uint32_t *s_wt = s_mut_dyn + TPB * MUT_STRIDE;  // wildtype, 16 sites per word
    const uint32_t L = st->L;
    const uint8_t *__restrict__ wt = st->wildtype;
    const uint32_t wt_words = mutate_wt_words(L);
    if (threadIdx.x < 16) s_subst[threadIdx.x] = st->host_subst[threadIdx.x];
    for (uint32_t k = threadIdx.x; k < wt_words; k += blockDim.x) {
        uint32_t word = 0;
        for (uint32_t q = 0; q < 16; q++) {
            const uint32_t site = k * 16 + q;
            if (site < L) word |= (uint32_t)wt[site] << (2 * q);
        }
        s_wt[k] = word;
    }
    __syncthreads();
    auto wt_at = [&](uint32_t site) -> uint32_t { return wt_words ? (s_wt[site >> 4] >> ((site & 15u) * 2u)) & 3u : (uint32_t)wt[site]; };
    const uint32_t w_first = range == 2 ? st->mut_own : 0u;
    const uint32_t n_list = (range == 1 ? st->mut_own : st->n_mut_list) - w_first;  // items of this launch
    const bool ahead = range == 1;  // (the arrays of the population are picked where they are used: nothing more to keep in registers)
    const uint32_t span = gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31;
    // Every warp walks the worklist 32 items at a time.  Same-address atomics are served one after the other by L2, so
    // none is spent per record or per 32 records: a record's id is its worklist position behind the record count the
    // kernel started with (published by the last CTA to leave), and arena words (needed only by the records that are
    // flattened) come out of a chunk the warp reserves for all the rounds it still has to do.
    // (BIGMODE 2 runs behind the main pass, which has already moved the record count past this launch's ids)
    const unsigned long long id_base = range == 1 ? st->mut_own_base : (BIGMODE == 2 ? st->n_hap - n_list : st->n_hap);
    const uint32_t n_items = BIGMODE == 2 ? st->n_big : n_list;
    unsigned long long chunk_off = 0;
    uint32_t chunk_left = 0;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_items; base += span) {
        bool ok = base + lane < n_items;
        const uint32_t w = BIGMODE == 2 ? (ok ? gptr(st->big_list)[base + lane] : 0u) : base + lane;
        uint32_t i = 0, nm = 0, pid = 0, tpos = NONE32;
        HapHead ph;
        ph.off = 0; ph.len = 0; ph.nd_mlen = 0; ph.raw0 = 1.0; ph.fit0 = 1.0;
        if (ok) {
            const uint4 item = gptr(st->mut_list)[w_first + w];
            i = item.x;
            nm = item.y;
            tpos = item.z;
            pid = parent_in_item ? item.w : gptr(ahead ? st->hap_id_next : st->hap_id)[i];
            if (BIGMODE == 1 && nm > MUT_LOCAL) {  // set aside for the second pass
                st->big_list[atomicAdd(&st->n_big, 1u)] = w;
                ok = false;
                nm = 0;
            } else {
                ph = gptr(st->head)[pid];  // ONE 64-byte line: base list, inline delta, fitness of the parent
            }
        }
        const uint32_t nd = head_nd(ph), plen = ph.len, mlen = head_mlen(ph);
        const bool flat = ok && nd + nm > HEAD_DELTA;  // the delta would overflow: this record gets a base list of its own
        const bool big = BIGMODE == 1 ? false : nm > MUT_LOCAL;  // (implies flat)
        if (flat) {  // the whole base list is about to be copied: pull it towards L2 while the sites are drawn
            const uint32_t *pl = st->arena + ph.off;
            for (uint32_t k = 0; k < plen; k += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pl + k));
        }
        const uint32_t words = flat ? mlen + (big ? 4u * nm + HEAD_DELTA : nm) : 0u;
        uint32_t inc = words;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        const uint32_t wtot = __shfl_sync(0xFFFFFFFFu, inc, 31);
        if (wtot > chunk_left) {  // (warp-uniform)
            const uint32_t rounds = (n_items - base + span - 1) / span;
            const unsigned long long want = (unsigned long long)wtot * rounds + ((unsigned long long)wtot * rounds >> 3) + 32u;
            chunk_left = (uint32_t)(want < 0xFFFFFFFFull ? want : (unsigned long long)wtot);
            if (lane == 0) chunk_off = atomicAdd(&st->arena_top, (unsigned long long)chunk_left);
            chunk_off = __shfl_sync(0xFFFFFFFFu, chunk_off, 0);
        }
        const unsigned long long abase = chunk_off;
        chunk_off += wtot;
        chunk_left -= wtot;
        const uint64_t roff = abase + (inc - words);
        const uint64_t id64 = id_base + w;
        if (ok && id64 >= st->cap_hap) { raise(st, DE_HAP_CAPACITY); ok = false; }
        if (ok && flat && roff + words > st->cap_arena) { raise(st, DE_ARENA_CAPACITY); ok = false; }
        const uint32_t new_id = (uint32_t)id64;
        const uint32_t *pe = gptr(st->arena) + ph.off;
        uint32_t *out = gptr(st->arena) + roff;  // (flat only) merged list [0, mlen + nm), then scratch of large events
        uint32_t *pd = s_scratch + threadIdx.x * MUT_STRIDE, *dl = pd + HEAD_DELTA, *ch_local = dl + 2 * HEAD_DELTA, *ne_local = ch_local + MUT_LOCAL;
        uint32_t *chs = big ? out + mlen + nm : ch_local;          // change words: pos << 8 | present << 7 | from << 4 | to
        uint32_t *nev = big ? out + mlen + 2 * nm : ne_local;      // the event's entries, one per distinct site
        uint32_t *dm = big ? out + mlen + 3 * nm : dl;             // parent delta merged with the event's entries
        uint32_t ndm = 0;
        int32_t dlen = 0;
        HapHead nh;
        nh.raw0 = 1.0;
        nh.fit0 = 1.0;
        double fit_repro = 1.0;  // mapped fitness under the reproductive provider of spec 0 (patch_records == 3)
        if (ok) {
#pragma unroll
            for (uint32_t k = 0; k < HEAD_DELTA; k++) pd[k] = ph.d[k];  // (indexed dynamically below: shared memory, not local memory)
            if (replay) {
                const uint64_t ro = rp.offsets[i];
                for (uint32_t k = 0; k < nm; k++) {
                    uint32_t site = rp.sites[ro + k];
                    if (site >= L || (k > 0 && site < ch_pos(chs[k - 1])) || rp.bases[ro + k] > 3) { raise(st, DE_REPLAY); site = site % L; }
                    const uint32_t e = merged_find(pe, plen, pd, nd, site, L);
                    const uint32_t cur = e != 0xFFFFFFFFu ? entry_g(e) : wt_at(site);
                    const uint32_t to = rp.bases[ro + k] & 3;
                    if (cur == to) raise(st, DE_SAME_BASE);
                    chs[k] = (site << 8) | (e != 0xFFFFFFFFu ? 0x80u : 0u) | (cur << 4) | to;
                }
            } else {
                Philox g(st->seed, st->compartment, st->generation + (ahead ? 1u : 0u), PH_MUTATE, i);
                for (uint32_t k = 0; k < nm; k++) chs[k] = (uint32_t)g.below(L);
                if (nm <= 16) {
                    for (uint32_t a = 1; a < nm; a++) {
                        uint32_t x = chs[a], c = a;
                        while (c > 0 && chs[c - 1] > x) { chs[c] = chs[c - 1]; c--; }
                        chs[c] = x;
                    }
                } else {
                    heap_sort_u32(chs, nm);
                }
                for (uint32_t k = 0; k < nm; k++) {
                    const uint32_t site = chs[k];
                    // Haplotype::get_base on the PRE-mutation haplotype (duplicate sites of one event see the same base)
                    const uint32_t e = merged_find(pe, plen, pd, nd, site, L);
                    const uint32_t cur = e != 0xFFFFFFFFu ? entry_g(e) : wt_at(site);
                    const uint32_t to = (uint32_t)weighted4(g, &s_subst[cur * 4]);
                    if (cur == to) raise(st, DE_SAME_BASE);
                    chs[k] = (site << 8) | (e != 0xFFFFFFFFu ? 0x80u : 0u) | (cur << 4) | to;
                }
            }
            // the event's entries: per distinct site g = first change's base (get_base: first matching change wins), m = last
            // change's base, absent when that is the wildtype base (get_mutations: last change wins, back-mutation removes the key)
            uint32_t n_new = 0;
            for (uint32_t b = 0; b < nm;) {
                const uint32_t pb = ch_pos(chs[b]), first_to = ch_to(chs[b]), present = ch_present(chs[b]);
                uint32_t last_to = first_to;
                while (b < nm && ch_pos(chs[b]) == pb) { last_to = ch_to(chs[b]); b++; }
                const uint32_t w0 = wt_at(pb);
                const uint32_t m = (last_to == w0) ? 4u : last_to;
                const bool gone = first_to == w0 && m == 4u;   // indistinguishable from the wildtype at this site
                if (gone && !present) continue;                 // nothing to override (a tombstone the parent's delta may hold stays)
                nev[n_new++] = gone ? make_entry(pb, w0, ENTRY_TOMBSTONE) : make_entry(pb, first_to, m);
                dlen += present ? (gone ? -1 : 0) : 1;
            }
            // parent delta merged with the event's entries (same position: the event wins)
            {
                uint32_t a = 0, b = 0;
                while (a < nd || b < n_new) {
                    const uint32_t pa = a < nd ? entry_pos(pd[a]) : 0xFFFFFFFFu, pb = b < n_new ? entry_pos(nev[b]) : 0xFFFFFFFFu;
                    if (pa < pb) dm[ndm++] = pd[a++];
                    else { if (pa == pb) a++; dm[ndm++] = nev[b++]; }
                }
            }
            const int pi_repro = st->specs[0].repro;
            for (uint32_t p = 0; p < st->n_providers; p++) {
                const DevProvider &pr = st->prov[p];
                double raw = 1.0;
                if (pr.kind != VL_FITNESS_NEUTRAL) {
                    const double praw = p == 0 ? ph.raw0 : gptr(st->h_raw[p])[pid];
                    double acc = 1.;
                    for (uint32_t k = 0; k < nm; k++) {
                        const uint64_t row = (uint64_t)ch_pos(chs[k]) * 4;
                        acc = acc * gptr(pr.table)[row + ch_to(chs[k])] / gptr(pr.table)[row + ch_from(chs[k])];
                    }
                    if (EPI && pr.kind == VL_FITNESS_EPISTATIC) acc = acc * epi_update(pr, chs, nm, pe, plen, dm, ndm, L);
                    raw = praw * acc;
                }
                const double mapped = utility_apply(pr, raw);
                gptr(st->h_raw[p])[new_id] = raw;
                gptr(st->h_fit[p])[new_id] = mapped;
                if (p == 0) { nh.raw0 = raw; nh.fit0 = mapped; }
                if ((int)p == pi_repro) fit_repro = mapped;
            }
        }
        // ---- records that get a list of their own: flattened by the whole warp, one after the other (the merged deltas
        // sit in shared memory — or, for events of more than MUT_LOCAL sites, in the arena — where every lane can read them)
        uint32_t flat_len = 0;
        {
            uint32_t jobs = __ballot_sync(0xFFFFFFFFu, ok && flat && !big);
            while (jobs) {
                const uint32_t src = __ffs(jobs) - 1;
                jobs &= jobs - 1;
                const uint64_t pe_s = __shfl_sync(0xFFFFFFFFu, (uint64_t)(uintptr_t)pe, src);
                const uint64_t dm_s = __shfl_sync(0xFFFFFFFFu, (uint64_t)(uintptr_t)dm, src);
                const uint64_t out_s = __shfl_sync(0xFFFFFFFFu, (uint64_t)(uintptr_t)out, src);
                const uint32_t plen_s = __shfl_sync(0xFFFFFFFFu, plen, src), ndm_s = __shfl_sync(0xFFFFFFFFu, ndm, src);
                const uint32_t o = warp_flatten((const uint32_t *)(uintptr_t)pe_s, plen_s, (const uint32_t *)(uintptr_t)dm_s, ndm_s, (uint32_t *)(uintptr_t)out_s, L);
                if (lane == src) flat_len = o;
            }
            if (ok && flat && big) {  // rare: more sites in one event than the shared scratch holds
                uint32_t o = 0;
                merged_for_each(pe, plen, dm, ndm, [&](uint32_t e) { out[o++] = e; });
                flat_len = o;
            }
        }
        if (!ok) continue;
        if (flat) {
            nh.off = roff;
            nh.len = flat_len;
            nh.nd_mlen = flat_len << 8;
#pragma unroll
            for (uint32_t k = 0; k < HEAD_DELTA; k++) nh.d[k] = 0xFFFFFFFFu;
        } else {
            nh.off = ph.off;
            nh.len = plen;
            nh.nd_mlen = ndm | ((uint32_t)((int32_t)mlen + dlen) << 8);
#pragma unroll
            for (uint32_t k = 0; k < HEAD_DELTA; k++) nh.d[k] = k < ndm ? dm[k] : 0xFFFFFFFFu;
        }
        gptr(st->head)[new_id] = nh;
        if (patch_records >= 3) {
            // device-resident generation (vl_fused.cuh): items carry the host in .z.  The new haplotype's mapped fitness
            // replaces the infectant's carried value and joins its host's sum (the pass that staged the item left it out).
            // 3: that pass has already written the new record's id into the population (ids are worklist positions);
            // 4: records were created between that pass and this kernel (imported migrants): the population is patched here.
            gptr(ahead ? st->pop_fit_next : st->pop_fit)[i] = fit_repro;
            if (tpos != NONE32 && fit_repro != 0.0) red_add_f64_hint(&st->hsum[ahead ? st->hsum_parity ^ 1u : st->hsum_parity][tpos], fit_repro, l2_keep_policy());
            if (patch_records == 4) gptr(ahead ? st->hap_id_next : st->hap_id)[i] = new_id;
            continue;
        }
        st->hap_id[i] = new_id;
        if (patch_records == 2) {
            if (tpos != NONE32) st->tmp_rec[tpos].y = new_id;  // the bucket-partitioned record, not yet sorted
        } else if (patch_records) {
            const uint32_t h = st->host_id[i];
            if (h != NONE32) {
                const uint32_t sb = st->offsets[h], se = st->offsets[h + 1];
                for (uint32_t q = sb; q < se; q++)
                    if (st->rec[q].x == i) { st->rec[q].y = new_id; break; }
            }
        }
    }
    if (BIGMODE != 2 && range == 1) return;  // (the ids were reserved up front)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&st->mut_ctas_done, 1u) == gridDim.x - 1) {  // the last CTA to leave
            if (BIGMODE == 2) {
                st->n_big = 0;  // (the list is done)
            } else {
                const unsigned long long top = id_base + n_list;
                st->n_hap = top < st->cap_hap ? top : st->cap_hap;
            }
            st->mut_ctas_done = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------ replication
// mapped fitness of a haplotype under the reproductive provider of spec s (or 1.0), src/simulation.rs:137-143
__device__ __forceinline__ double reproductive_fitness(const DevState *st, int s, uint32_t hap) {
    const int pi = s >= 0 ? st->specs[s].repro : -1;
    return pi >= 0 ? st->h_fit[pi][hap] : 1.;
}
// lambda as the offspring kernel consumes it: 0.0 stands for "no offspring".  Poisson::new fails for a mean that
// is not finite and > 0; the reference then leaves the f64 bit pattern of f in the slot (src/simulation.rs:146-152),
// which is 0 offspring for f == 0.0 and undefined dynamics otherwise — reported, not imitated.
__device__ __forceinline__ double checked_lambda(DevState *st, double f, double R0, double S) {
    const double lam = f * R0 / S;
    if (isfinite(lam) && lam > 0.0 && lam <= 1.844e19) return lam;
    if (__double_as_longlong(f) != 0ll) raise(st, DE_UNDEFINED_OFFSPRING);
    return 0.0;
}

// spec of a CSR position: host ranges are contiguous, so each spec owns the position range
// [offsets[lo], offsets[min(hi, H)]); first match wins like HostSpecs::try_get_spec_from_index
__device__ __forceinline__ int find_spec_of_position(const DevState *st, uint64_t p) {
    for (uint32_t s = 0; s < st->n_specs; s++) {
        const uint64_t lo = st->specs[s].lo < st->H ? st->specs[s].lo : st->H, hi = st->specs[s].hi < st->H ? st->specs[s].hi : st->H;
        if (lo < hi && p >= st->offsets[lo] && p < st->offsets[hi]) return (int)s;
    }
    return -1;
}

// One thread per CSR position: mapped reproductive fitness of the record's haplotype (the random gather of the
// generation), parked in lam[p] until k_host_sum turns it into the Poisson mean.
// `guarded`: run only if the fused replicate kernel gave up on this population (DevState::fused_failed)
__global__ void __launch_bounds__(TPB) k_fitness(DevState *st, int guarded) {
    if (guarded && !st->fused_failed) return;
    const uint64_t m = st->n_sorted;
    const uint2 *__restrict__ rec = st->rec;
    double *__restrict__ lam = st->lam;
    const bool single = st->n_specs == 1 && st->specs[0].lo == 0 && st->specs[0].hi >= st->H;
    const int pi0 = st->specs[0].repro;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (uint64_t)gridDim.x * blockDim.x) {
        double f;
        if (single) f = pi0 >= 0 ? st->h_fit[pi0][rec[p].y] : 1.;
        else f = reproductive_fitness(st, find_spec_of_position(st, p), rec[p].y);
        lam[p] = f;
    }
}

// sort one slice by descending infectant index, carrying the parked fitness values along
__device__ __forceinline__ void sort_slice_desc(uint2 *rec, double *lam, uint32_t b, uint32_t e) {
    for (uint32_t a = b + 1; a < e; a++) {
        const uint2 x = rec[a];
        const double fx = lam[a];
        uint32_t c = a;
        while (c > b && rec[c - 1].x < x.x) { rec[c] = rec[c - 1]; lam[c] = lam[c - 1]; c--; }
        rec[c] = x;
        lam[c] = fx;
    }
}

// One thread per host: order the slice, S_h = sum of f over the slice in slice order from 0.0 (iter().sum(),
// src/simulation.rs:144), left at every member's CSR position (hs[p]); the division that turns (f, S) into the
// Poisson mean happens position-parallel in k_offspring.  Long slices go to the heavy list.
__global__ void __launch_bounds__(TPB) k_host_sum(DevState *st, int order_only, int guarded) {
    if (guarded && !st->fused_failed) return;
    const uint64_t H = st->H;
    const uint32_t *__restrict__ offsets = st->offsets;
    uint2 *rec = st->rec;
    double *lam = st->lam;
    double *hs = st->hs;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < H; h += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = offsets[h], e = offsets[h + 1];
        const uint32_t len = e - b;
        if (len == 0) continue;
        if (len <= 4) {
            // up to four members in registers, all loads issued at once; absent members get the lowest key and
            // f = +0.0 (adding +0.0 to a non-negative partial sum is exact), five compare-exchanges, larger index first
            uint2 r0 = rec[b], r1 = len > 1 ? rec[b + 1] : make_uint2(0u, 0u), r2 = len > 2 ? rec[b + 2] : make_uint2(0u, 0u),
                  r3 = len > 3 ? rec[b + 3] : make_uint2(0u, 0u);
            double f0 = lam[b], f1 = len > 1 ? lam[b + 1] : 0.0, f2 = len > 2 ? lam[b + 2] : 0.0, f3 = len > 3 ? lam[b + 3] : 0.0;
            if (len > 1) {
                uint32_t k0 = r0.x + 1u, k1 = len > 1 ? r1.x + 1u : 0u, k2 = len > 2 ? r2.x + 1u : 0u, k3 = len > 3 ? r3.x + 1u : 0u;
#define VL_CX(ka, ra, fa, kb, rb, fb) if (ka < kb) { const uint32_t tk = ka; ka = kb; kb = tk; const uint2 tr = ra; ra = rb; rb = tr; const double tf = fa; fa = fb; fb = tf; }
                VL_CX(k0, r0, f0, k1, r1, f1) VL_CX(k2, r2, f2, k3, r3, f3) VL_CX(k0, r0, f0, k2, r2, f2) VL_CX(k1, r1, f1, k3, r3, f3) VL_CX(k1, r1, f1, k2, r2, f2)
#undef VL_CX
                rec[b] = r0; lam[b] = f0;
                rec[b + 1] = r1; lam[b + 1] = f1;
                if (len > 2) { rec[b + 2] = r2; lam[b + 2] = f2; }
                if (len > 3) { rec[b + 3] = r3; lam[b + 3] = f3; }
            }
            if (!order_only) {
                const double S = (((0.0 + f0) + f1) + f2) + f3;
                hs[b] = S;
                if (len > 1) hs[b + 1] = S;
                if (len > 2) hs[b + 2] = S;
                if (len > 3) hs[b + 3] = S;
            }
            continue;
        }
        if (len > THREAD_TIER) {
            const uint32_t slot = atomicAdd(&st->n_heavy, 1u);
            st->heavy_hosts[slot] = (uint32_t)h;
            continue;
        }
        // 5 .. THREAD_TIER members: rare at N ~ H, so they are compacted into a worklist instead of stalling the warp
        st->medium_hosts[atomicAdd(&st->n_medium, 1u)] = (uint32_t)h;
    }
}
// One thread per listed host (5 .. THREAD_TIER members): insertion sort, sequential sum
__global__ void __launch_bounds__(TPB) k_host_sum_medium(DevState *st, int order_only, int guarded) {
    if (guarded && !st->fused_failed) return;
    const uint32_t n_medium = st->n_medium;
    const uint32_t *__restrict__ offsets = st->offsets;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_medium; w += gridDim.x * blockDim.x) {
        const uint32_t h = st->medium_hosts[w];
        const uint32_t b = offsets[h], e = offsets[h + 1];
        sort_slice_desc(st->rec, st->lam, b, e);
        if (order_only) continue;
        double S = 0.0;
        for (uint32_t a = b; a < e; a++) S += st->lam[a];
        for (uint32_t a = b; a < e; a++) st->hs[a] = S;
    }
}

// One CTA per heavy slice: normalised bitonic network on (index << 32 | haplotype) keys with the fitness as
// payload, larger index first (missing partners beyond the slice end are no-ops), in shared memory when the
// slice fits, else in global memory; then the sum strictly sequential in slice order, lambda in parallel.
__global__ void __launch_bounds__(TPB) k_host_sum_heavy(DevState *st, int order_only, int guarded) {
    if (guarded && !st->fused_failed) return;
    extern __shared__ unsigned long long s_dyn[];
    __shared__ double s_S;
    unsigned long long *s_keys = s_dyn;
    double *s_f = reinterpret_cast<double *>(s_dyn + SMEM_SORT_MAX);
    const uint32_t n_heavy = st->n_heavy;
    for (uint32_t w = blockIdx.x; w < n_heavy; w += gridDim.x) {
        const uint32_t h = st->heavy_hosts[w];
        const uint32_t b = st->offsets[h], len = st->offsets[h + 1] - b;
        unsigned long long *g = reinterpret_cast<unsigned long long *>(st->rec + b);  // uint2 {x = index, y = hap}: as u64 = y << 32 | x
        double *gf = st->lam + b;
        const bool in_smem = len <= SMEM_SORT_MAX;
        unsigned long long *a = in_smem ? s_keys : g;
        double *af = in_smem ? s_f : gf;
        // key = index << 32 | hap so that comparing keys compares indices
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
            const unsigned long long v = g[i];
            a[i] = (v << 32) | (v >> 32);
            if (in_smem) af[i] = gf[i];
        }
        __syncthreads();
        uint32_t P = 1;
        while (P < len) P <<= 1;
        for (uint32_t k = 2; k <= P; k <<= 1) {
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
                uint32_t l = i ^ (k - 1);
                if (l > i && l < len) {
                    unsigned long long x = a[i], y = a[l];
                    if (x < y) { a[i] = y; a[l] = x; const double t = af[i]; af[i] = af[l]; af[l] = t; }
                }
            }
            __syncthreads();
            for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
                    uint32_t l = i ^ j;
                    if (l > i && l < len) {
                        unsigned long long x = a[i], y = a[l];
                        if (x < y) { a[i] = y; a[l] = x; const double t = af[i]; af[i] = af[l]; af[l] = t; }
                    }
                }
                __syncthreads();
            }
        }
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
            const unsigned long long v = a[i];
            g[i] = (v << 32) | (v >> 32);
            if (in_smem) gf[i] = af[i];
        }
        __syncthreads();
        if (order_only) continue;
        if (threadIdx.x < 32) {  // lanes fetch 32 terms at a time, the additions stay strictly sequential
            const uint32_t lane = threadIdx.x;
            double S = 0.0;
            for (uint32_t a0 = 0; a0 < len; a0 += 32) {
                const double f = (a0 + lane < len) ? gf[a0 + lane] : 0.0;
                const uint32_t cnt = min(32u, len - a0);
                for (uint32_t t = 0; t < cnt; t++) S += __shfl_sync(0xFFFFFFFFu, f, t);
            }
            if (lane == 0) s_S = S;
        }
        __syncthreads();
        const double S = s_S;
        double *ghs = st->hs + b;
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) ghs[i] = S;
        __syncthreads();
    }
}

// offspring of the record at CSR position p ~ Poisson(lam[p]) (src/simulation.rs:148-152); one CTA per resample
// tile so that the tile's offspring total comes out of the same pass.  replay_off is indexed by infectant.
// SMALL_ONLY: every mean is below 12 (R0 < 12 and lambda = R0 * f/S <= R0), so the rejection sampler is not compiled in.
// store_lambda: keep the Poisson mean in lam[p] for vl_get_replication_terms (trait-level calls; the fused loop skips the store)
template <bool SMALL_ONLY>
__global__ void __launch_bounds__(TPB) k_offspring(DevState *st, const uint64_t *__restrict__ replay_off, int store_lambda, int guarded) {
    __shared__ uint64_t s_red[TPB / 32];
    __shared__ double s_rcp[POISSON_RCP];
    if (guarded && !st->fused_failed) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->host_tiles = 0;  // this kernel's tile totals are per RESAMPLE_TILE positions
    poisson_rcp_init(s_rcp);
    __syncthreads();
    const uint64_t m = st->n_sorted;
    const uint64_t n_tiles = (m + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
    double *__restrict__ lam = st->lam;
    const double *__restrict__ hs = st->hs;
    const uint2 *__restrict__ rec = st->rec;
    uint32_t *__restrict__ offspring = st->offspring;
    const uint32_t gen = st->generation;
    const double R0 = st->host_r0;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t local = 0;
#pragma unroll 2
        for (uint32_t k = 0; k < RESAMPLE_TILE / TPB; k++) {
            const uint64_t p = tile * RESAMPLE_TILE + (uint64_t)k * TPB + threadIdx.x;
            if (p >= m) break;
            uint32_t off = 0;
            const double l = checked_lambda(st, lam[p], R0, hs[p]);  // lambda_i = f_i * R0 / S_host (src/simulation.rs:148-150)
            if (store_lambda) lam[p] = l;
            if (replay_off) {
                uint64_t v = replay_off[rec[p].x];
                if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0; }
                off = (uint32_t)v;
            } else {
                if (l > 0.0) {
                    if (SMALL_ONLY && !(l < 12.0)) raise(st, DE_UNDEFINED_OFFSPRING);  // f > S: negative fitness values in the slice
                    // the stream of a draw is addressed by the infectant, not by where its record sits: every layout of
                    // the host map (dense CSR positions here, tile-strided in vl_tiles.cuh) draws the same count
                    uint64_t v = poisson_draw<SMALL_ONLY>(key, rec[p].x, gen, PH_OFFSPRING, l, s_rcp);
                    if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0xFFFFFFFFull; }
                    off = (uint32_t)v;
                }
            }
            offspring[p] = off;
            local += off;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, s);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t t = 0;
            for (int w = 0; w < TPB / 32; w++) t += s_red[w];
            st->tile_sum[tile] = t;
        }
        __syncthreads();
    }
}
// ---- fused replicate: fitness gather + in-host order + per-host sum + lambda + Poisson in ONE pass -----------------
// One CTA per host tile (`ht` consecutive hosts = a contiguous range of CSR positions, whole hosts only), every phase
// position-parallel with the tile's records, fitness values and host boundaries in shared memory:
//   1. stage offsets and records (coalesced), gather the mapped fitness of every record (the generation's random
//      gather), mark the first position of every non-empty host in a bit mask;
//   2. every record finds its slice in the bit mask, ranks itself by descending infectant index (the reference's
//      in-host order) and moves to its rank (written back to rec[] only where that changes something);
//   3. every position sums its slice in slice order from 0.0, forms lambda = f * R0 / S, draws Poisson(lambda),
//      stores the offspring count (coalesced) and adds to the tile total.
// Same arithmetic, same Philox addresses and the same results as k_fitness + k_host_sum* + k_offspring; what it saves
// is three round trips of per-position f64 arrays through HBM and the host-parallel passes' idle lanes.  A tile that
// does not fit (more than RESAMPLE_TILE records, or a slice beyond the thread tier) sets DevState::fused_failed and
// the general kernels redo the phase (`fail_is_error`: the caller enqueued no fallback because uniform infection
// makes that a > 8 sigma event).
constexpr int REPL_TPB = 512;
constexpr int REPL_PER = HOST_TILE_CAP / REPL_TPB;
template <bool SMALL_ONLY>
__global__ void __launch_bounds__(REPL_TPB, 3) k_replicate_fused(DevState *st, const uint64_t *__restrict__ replay_off, int store_lambda,
                                                                 int fail_is_error, uint32_t tiles_per_cta) {
    __shared__ uint32_t s_off[HT_MAX_HOSTS + 1];
    __shared__ uint32_t s_head[HOST_TILE_CAP / 32 + 2];
    __shared__ double s_f[HOST_TILE_CAP];
    __shared__ uint2 s_rec[HOST_TILE_CAP];
    __shared__ uint32_t s_red[REPL_TPB / 32];
    const uint64_t H = st->H;
    const uint32_t ht = st->ht_planned;  // host_tile_size(n, H), from the state, not from the launch (k_plan_host_tiles)
    if (ht == 0) {  // more infectants per host than whole-host tiles can stage
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->fused_failed = 1u; if (fail_is_error) raise(st, DE_TILE_OVERFLOW); }
        return;
    }
    const uint64_t nt = (H + ht - 1) / ht;
    const uint32_t *__restrict__ offsets = st->offsets;
    uint2 *rec = st->rec;
    uint32_t *__restrict__ offspring = st->offspring;
    double *__restrict__ lam = st->lam;
    const uint32_t gen = st->generation;
    const double R0 = st->host_r0;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const bool single = st->n_specs == 1 && st->specs[0].lo == 0 && st->specs[0].hi >= H;
    const int pi0 = st->specs[0].repro;
    const uint32_t tid = threadIdx.x;
    const uint32_t kmax = poisson_kmax(R0 < 12.0 ? R0 : 12.0);  // lambda = R0 * f / S <= R0 for non-negative fitness
    const uint64_t n_sorted = st->n_sorted;
    if (blockIdx.x == 0 && tid == 0) st->host_tiles = ht;
    // a CTA takes consecutive tiles, so while it works on one it knows where the next one's offsets and records are
    const uint64_t tile_end = (uint64_t)(blockIdx.x + 1) * tiles_per_cta < nt ? (uint64_t)(blockIdx.x + 1) * tiles_per_cta : nt;
    for (uint64_t tile = (uint64_t)blockIdx.x * tiles_per_cta; tile < tile_end; tile++) {
        const uint64_t h0 = tile * ht;
        const uint32_t nh = (uint32_t)(H - h0 < ht ? H - h0 : ht);
        __syncthreads();  // shared arrays of the previous tile are free
        for (uint32_t k = tid; k <= nh; k += REPL_TPB) s_off[k] = offsets[h0 + k];
        for (uint32_t k = tid; k < HOST_TILE_CAP / 32 + 2; k += REPL_TPB) s_head[k] = 0u;
        __syncthreads();
        const uint32_t p0 = s_off[0], cnt = s_off[nh] - p0;
        if (cnt > HOST_TILE_CAP) {  // uniform across the CTA
            if (tid == 0) { st->fused_failed = 1u; st->tile_sum[tile] = 0; if (fail_is_error) raise(st, DE_TILE_OVERFLOW); }
            continue;
        }
        // ---- 1. records and fitness into registers and shared memory (slot j of thread t is position t + j * TPB)
        uint2 r[REPL_PER];
        double f[REPL_PER];
#pragma unroll
        for (int j = 0; j < REPL_PER; j++) {
            const uint32_t q = tid + j * REPL_TPB;
            r[j] = q < cnt ? rec[p0 + q] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < REPL_PER; j++) {
            const uint32_t q = tid + j * REPL_TPB;
            f[j] = 0.0;
            if (q < cnt) {
                if (single) f[j] = pi0 >= 0 ? st->h_fit[pi0][r[j].y] : 1.;
                else f[j] = reproductive_fitness(st, find_spec_of_position(st, (uint64_t)p0 + q), r[j].y);
            }
        }
#pragma unroll
        for (int j = 0; j < REPL_PER; j++) {
            const uint32_t q = tid + j * REPL_TPB;
            if (q < cnt) {
                s_rec[q] = r[j];
                s_f[q] = f[j];
            }
        }
        if (tile + 1 < tile_end) {  // pull the next tile's offsets and records towards L2 while this one is computed
            const uint64_t hn = h0 + ht + (uint64_t)tid * 8;
            if (hn <= H) asm volatile("prefetch.global.L2 [%0];" ::"l"(offsets + hn));
            if ((uint64_t)p0 + cnt + (uint64_t)tid * 4 < n_sorted) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + (uint64_t)p0 + cnt + (uint64_t)tid * 4));
        }
        bool failed = false;
        for (uint32_t k = tid; k < nh; k += REPL_TPB) {
            const uint32_t b = s_off[k] - p0, len = s_off[k + 1] - s_off[k];
            if (len) atomicOr(&s_head[b >> 5], 1u << (b & 31));
            if (len > THREAD_TIER) failed = true;
        }
        if (tid == 0) atomicOr(&s_head[cnt >> 5], 1u << (cnt & 31));  // sentinel: the position after the last slice
        if (__syncthreads_or(failed ? 1 : 0)) {
            if (tid == 0) { st->fused_failed = 1u; st->tile_sum[tile] = 0; if (fail_is_error) raise(st, DE_TILE_OVERFLOW); }
            continue;
        }
        // ---- 2. slice bounds from the bit mask (a slice spans at most two words), rank by descending infectant index
        uint32_t seg[REPL_PER];  // first position of the slice | length << 16
        uint32_t dst[REPL_PER];
#pragma unroll
        for (int j = 0; j < REPL_PER; j++) {
            const uint32_t q = tid + j * REPL_TPB;
            seg[j] = 0;
            dst[j] = q;
            if (q >= cnt) continue;
            const uint32_t w = q >> 5, bit = q & 31;
            const uint32_t below = s_head[w] & (0xFFFFFFFFu >> (31 - bit));
            const uint32_t b = below ? (w << 5) + 31 - __clz(below) : ((w - 1) << 5) + 31 - __clz(s_head[w - 1]);
            const uint32_t above = bit == 31 ? 0u : s_head[w] & (0xFFFFFFFFu << (bit + 1));
            const uint32_t e = above ? (w << 5) + __ffs(above) - 1 : ((w + 1) << 5) + __ffs(s_head[w + 1]) - 1;
            uint32_t rank = 0;
            for (uint32_t a = b; a < e; a++) rank += s_rec[a].x > r[j].x ? 1u : 0u;
            seg[j] = b | ((e - b) << 16);
            dst[j] = b + rank;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < REPL_PER; j++) {
            const uint32_t q = tid + j * REPL_TPB;
            if (q < cnt && dst[j] != q) {
                s_rec[dst[j]] = r[j];
                s_f[dst[j]] = f[j];
                rec[p0 + dst[j]] = r[j];
            }
        }
        __syncthreads();
        // ---- 3. per position: S in slice order, lambda, Poisson(lambda), offspring, tile total
        uint32_t local = 0;
#pragma unroll
        for (int j = 0; j < REPL_PER; j++) {
            const uint32_t q = tid + j * REPL_TPB;
            if (q >= cnt) continue;
            const uint32_t b = seg[j] & 0xFFFFu, e = b + (seg[j] >> 16);
            double S = 0.0;
            for (uint32_t a = b; a < e; a++) S += s_f[a];
            const double l = checked_lambda(st, s_f[q], R0, S);  // lambda_i = f_i * R0 / S_host (src/simulation.rs:148-150)
            const uint64_t p = (uint64_t)p0 + q;
            uint32_t off = 0;
            if (replay_off) {
                uint64_t v = replay_off[s_rec[q].x];
                if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0; }
                off = (uint32_t)v;
            } else if (l > 0.0) {
                uint64_t v;
                const uint64_t who = s_rec[q].x;  // (streams are addressed by the infectant: see k_offspring)
                if (l < 12.0) v = poisson_draw_small_fast(key, who, gen, PH_OFFSPRING, l, c_rcp64, kmax)  /* the f64 fallback reads 1/k from constant memory */;
                else if (SMALL_ONLY) { raise(st, DE_UNDEFINED_OFFSPRING); v = 0; }  // f > S: negative fitness values in the slice
                else v = poisson_ptrs(key, who, gen, PH_OFFSPRING, l);
                if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0xFFFFFFFFull; }
                off = (uint32_t)v;
            }
            offspring[p] = off;
            if (store_lambda) lam[p] = l;
            local += off;
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, sh);
        if ((tid & 31) == 0) s_red[tid >> 5] = local;
        __syncthreads();
        if (tid == 0) {
            uint64_t t = 0;
            for (int w = 0; w < REPL_TPB / 32; w++) t += s_red[w];
            st->tile_sum[tile] = t;
        }
    }
}

// tile sums of an offspring map that came from the host (identity layout)
__global__ void __launch_bounds__(TPB) k_tile_sums(DevState *st) {
    __shared__ uint64_t s_red[TPB / 32];
    if (blockIdx.x == 0 && threadIdx.x == 0) st->host_tiles = 0;
    const uint64_t m = st->n_sorted;
    const uint64_t n_tiles = (m + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t local = 0;
        for (uint32_t k = 0; k < RESAMPLE_TILE / TPB; k++) {
            const uint64_t i = tile * RESAMPLE_TILE + (uint64_t)k * TPB + threadIdx.x;
            if (i < m) local += st->offspring[i];
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, s);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t t = 0;
            for (int w = 0; w < TPB / 32; w++) t += s_red[w];
            st->tile_sum[tile] = t;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------ resample
__device__ __forceinline__ uint64_t f64_as_usize(double x) {  // Rust `as usize`: saturating, NaN -> 0
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551615.0) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)x;
}
__device__ __forceinline__ uint64_t pow2_ceil(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

constexpr int PLAN_THREADS = 1024;

// block-wide sum of one u64 per thread (PLAN_THREADS threads)
__device__ __forceinline__ uint64_t plan_block_sum(uint64_t x, uint64_t *s_scratch) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_scratch[threadIdx.x >> 5] = x;
    __syncthreads();
    uint64_t t = 0;
    for (int w = 0; w < PLAN_THREADS / 32; w++) t += s_scratch[w];
    return t;
}
// exclusive scan of src[0..nt) into dst (may alias); returns nothing, total left in *s_carry
__device__ __forceinline__ void plan_block_scan(const uint64_t *src, uint64_t *dst, uint64_t nt, bool inclusive, uint64_t init, uint64_t *s_scan, uint64_t *s_carry) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __syncthreads();
    if (tid == 0) *s_carry = init;
    __syncthreads();
    constexpr uint32_t E = 16;  // up to E consecutive elements per thread in ONE round: one latency, one block scan
    if (nt <= (uint64_t)PLAN_THREADS * E) {
        const uint64_t per = (nt + PLAN_THREADS - 1) / PLAN_THREADS;
        const uint64_t lo = per * tid < nt ? per * tid : nt, hi = lo + per < nt ? lo + per : nt;
        uint64_t x[E], sum = 0;
#pragma unroll
        for (uint32_t e = 0; e < E; e++) { x[e] = lo + e < hi ? __ldcg(&src[lo + e]) : 0; sum += x[e]; }
        uint64_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        if (lane == 31) s_scan[wid] = inc;
        __syncthreads();
        uint64_t woff = 0, tot = 0;
        for (int w = 0; w < PLAN_THREADS / 32; w++) {
            uint64_t v = s_scan[w];
            if (w < (int)wid) woff += v;
            tot += v;
        }
        uint64_t run = init + woff + inc - sum;
#pragma unroll
        for (uint32_t e = 0; e < E; e++) {
            if (lo + e < hi) { dst[lo + e] = inclusive ? run + x[e] : run; run += x[e]; }
        }
        __syncthreads();
        if (tid == 0) *s_carry = init + tot;
        __syncthreads();
        return;
    }
    for (uint64_t t0 = 0; t0 < nt; t0 += PLAN_THREADS) {
        const uint64_t t = t0 + tid;
        const uint64_t x = t < nt ? __ldcg(&src[t]) : 0;  // (other CTAs may have added to it: read past L1)
        uint64_t inc = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        if (lane == 31) s_scan[wid] = inc;
        __syncthreads();
        uint64_t woff = 0, tot = 0;
        for (int w = 0; w < PLAN_THREADS / 32; w++) {
            uint64_t v = s_scan[w];
            if (w < (int)wid) woff += v;
            tot += v;
        }
        const uint64_t carry = *s_carry;
        if (t < nt) dst[t] = carry + woff + inc - (inclusive ? 0 : x);
        __syncthreads();
        if (tid == 0) *s_carry = carry + tot;
        __syncthreads();
    }
}

// PLAN_CTAS co-resident CTAs (far fewer than SMs) with a grid barrier on a counter in DevState.
// (1) Every CTA scans the tile totals for itself: W = sum, target_size = min(W*dilution, max_population)
// (src/simulation.rs:491-503) and size = (factor*target) as usize (:521) or an explicit amount (sample_indices, :505-514).
// (2) The `size` iid draws proportional to offspring (Population::sample, src/core/population.rs:385-403) are split
// over the tiles as an exact multinomial, every CTA drawing for its own slice of tiles.  Default: Poissonisation —
// independent N_t ~ Poisson(mu * w_t / W) with mu a few standard deviations below `size` are, given their sum K,
// Multinomial(K, w/W); the deficit size - K is Poissonised again (same argument, smaller means) until a handful of
// thousand draws remain, which are made one by one (binary search in the tile prefix, spread over all plan threads) and
// added: a multinomial plus an independent multinomial with the same cell probabilities is the multinomial of the summed size.  K > size
// (probability ~3e-7 per call) or `force_tree` falls back to walking a segment tree with conditional binomials on
// CTA 0, which is also exact.  (3) exclusive scan of the per-tile draw counts, slice by slice.
constexpr int PLAN_CTAS = 16;
constexpr uint64_t PLAN_FILL_MAX = 32768;  // a deficit up to this many draws is filled draw by draw (two per plan thread), no further stage
__device__ __forceinline__ void plan_grid_barrier(DevState *st, uint32_t &epoch) {
    __syncthreads();
    epoch += 1;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&st->plan_sync, 1u);
        const uint32_t target = epoch * gridDim.x;
        uint32_t seen;
        // The grid is at most 2 * PLAN_CTAS CTAs, far below what one B200 holds, but residency is not a guarantee (another
        // stream may own the SMs for a while): a barrier nobody else reaches within two seconds is reported, not waited out.
        const unsigned long long t0 = vl_globaltimer();
        uint32_t polls = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&st->plan_sync) : "memory");
            if (seen < target && (++polls & 1023u) == 0u && vl_globaltimer() - t0 > 2000000000ull) { raise(st, DE_PLAN_TIMEOUT); break; }
        } while (seen < target);
    }
    __syncthreads();
}
__device__ __forceinline__ uint64_t plan_ld(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// what to draw: n_parts subsamples of the same offspring map (the per-target parts of one origin, src/runner.rs:404-416,
// or the single part of next_generation); every part has its own Philox stream (salt)
struct PlanArgs {
    uint32_t n_parts;
    int use_amount, set_next, force_tree;
    double factor[MAX_PARTS];
    uint64_t amount[MAX_PARTS];
    uint32_t salt[MAX_PARTS];
};
__global__ void __launch_bounds__(PLAN_THREADS) k_plan_resample(DevState *st, PlanArgs pa, uint32_t smem_tiles) {
    extern __shared__ uint64_t s_prefix[];  // inclusive prefix of tile totals when nt <= smem_tiles
    __shared__ uint64_t s_scan[PLAN_THREADS / 32];
    __shared__ uint64_t s_carry;
    __shared__ uint64_t s_size[MAX_PARTS], s_W;
    __shared__ unsigned long long s_acc[MAX_PARTS];
    __shared__ double s_rcp[POISSON_RCP];
    poisson_rcp_init(s_rcp);
    const bool lead = blockIdx.x == 0;
    if (lead && threadIdx.x == 0) st->dbg[0] = vl_globaltimer();
    const uint64_t n = st->n;
    const uint64_t nt = tile_count(st);
    const uint64_t stride = st->tile_stride;
    const uint32_t tid = threadIdx.x;
    const uint32_t NP = pa.n_parts;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const uint32_t gen = st->generation;
    uint32_t epoch = 0;
    // leaving the kernel: the last CTA (every CTA has then passed every barrier) leaves the barrier counter and the stage
    // accumulators cleared for the next launch — no clearing pass in front of the kernel
    auto plan_leave = [&]() {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&st->plan_exit, 1u) == gridDim.x - 1) {
                for (int s = 0; s < 4; s++)
                    for (int p = 0; p < MAX_PARTS; p++) st->plan_acc[s][p] = 0ull;
                st->plan_exit = 0u;
                __threadfence();
                st->plan_sync = 0u;
            }
        }
    };
    // this CTA's slice of tiles
    const uint64_t per = (nt + gridDim.x - 1) / gridDim.x;
    const uint64_t t_lo = per * blockIdx.x < nt ? per * blockIdx.x : nt, t_hi = t_lo + per < nt ? t_lo + per : nt;
    uint64_t *prefix = nt <= smem_tiles ? s_prefix : st->tile_tree;  // (global fallback: every CTA writes the same values)
    plan_block_scan(st->tile_sum, prefix, nt, true, 0, s_scan, &s_carry);
    if (tid < NP) {
        const uint64_t total = s_carry;
        double target = 0.0;
        if (n > 0) {
            double a = (double)total * st->dilution, b = (double)st->max_population;
            target = a <= b ? a : b;
        }
        uint64_t size = pa.use_amount == 2 ? st->plan_amount[tid] : pa.use_amount ? pa.amount[tid] : f64_as_usize(pa.factor[tid] * target);
        if (n == 0) size = 0;
        if (n > 0 && total == 0) { if (lead && !(pa.use_amount == 2 && size == 0)) raise(st, DE_ALL_ZERO_WEIGHTS); size = 0; }
        if (pa.set_next && size > st->cap_inf) { if (lead) raise(st, DE_INF_CAPACITY); size = st->cap_inf; }
        if (lead) {
            if (tid == 0) {
                st->offspring_total = total;
                st->target_size = target;
                st->n_next = size;
            }
            st->plan_size[tid] = size;
        }
        s_size[tid] = size;
        if (tid == 0) s_W = total;
    }
    __syncthreads();
    const uint64_t W = s_W;
    if (lead && tid == 0) st->dbg[1] = vl_globaltimer();
    uint32_t tree_mask = pa.force_tree ? (NP >= 32 ? 0xFFFFFFFFu : (1u << NP) - 1u) : 0u;
    uint64_t K[MAX_PARTS];
    // Every stage below works on all (part, tile) pairs of this CTA's slice at once: a warp takes 32 consecutive tiles of
    // one part, so eight parts cost what one does instead of eight passes with a third of the threads busy.
    const uint32_t lane = tid & 31, wid = tid >> 5;
    const uint32_t chunks = (uint32_t)((t_hi - t_lo + 31) / 32);  // warp-sized runs of tiles in this CTA's slice
    if (!pa.force_tree) {
        if (tid < (uint32_t)MAX_PARTS) s_acc[tid] = 0ull;
        __syncthreads();
        for (uint32_t ch = wid; ch < NP * chunks; ch += PLAN_THREADS / 32) {  // stage 0
            const uint32_t p = ch / chunks;
            const uint64_t t = t_lo + (uint64_t)(ch - p * chunks) * 32 + lane;
            const uint64_t size = s_size[p];
            double mu = (double)size - 5.0 * sqrt((double)size) - 8.0;
            if (!(mu > 0.0) || size <= PLAN_FILL_MAX) mu = 0.0;  // a small sample is drawn one by one below, no Poissonisation needed
            uint64_t c = 0;
            if (t < t_hi) {
                const uint64_t w = st->tile_sum[t];
                if (w > 0 && mu > 0.0) c = poisson_draw(key, t | ((uint64_t)pa.salt[p] << 48), gen, PH_PLAN, mu * ((double)w / (double)W), s_rcp);
                st->tile_cnt[p * stride + t] = c;
            }
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, sh);
            if (lane == 0 && c) atomicAdd(&s_acc[p], (unsigned long long)c);
        }
        __syncthreads();
        if (tid < NP && s_acc[tid]) atomicAdd((unsigned long long *)&st->plan_acc[0][tid], s_acc[tid]);
        plan_grid_barrier(st, epoch);
        for (uint32_t p = 0; p < NP; p++) K[p] = plan_ld(&st->plan_acc[0][p]);
        if (lead && tid == 0) st->dbg[2] = vl_globaltimer();
        // the deficit size - K of every part is itself Poissonised (same argument, smaller means) until a handful of draws remain
        for (uint32_t stage = 1; stage < 4; stage++) {
            uint32_t active = 0;
            for (uint32_t p = 0; p < NP; p++) {
                const uint64_t size = s_size[p];
                if (K[p] <= size && size - K[p] > PLAN_FILL_MAX) {
                    const double d = (double)(size - K[p]);
                    if (d - 5.0 * sqrt(d) - 8.0 > 0.0) active |= 1u << p;
                }
            }
            if (!active) break;  // the same decision in every CTA: K comes from the shared accumulators
            if (tid < (uint32_t)MAX_PARTS) s_acc[tid] = 0ull;
            __syncthreads();
            for (uint32_t ch = wid; ch < NP * chunks; ch += PLAN_THREADS / 32) {
                const uint32_t p = ch / chunks;
                if (!(active >> p & 1u)) continue;  // (warp-uniform)
                const uint64_t t = t_lo + (uint64_t)(ch - p * chunks) * 32 + lane;
                const double d = (double)(s_size[p] - K[p]);
                const double mu2 = d - 5.0 * sqrt(d) - 8.0;
                uint64_t c = 0;
                if (t < t_hi) {
                    const uint64_t w = st->tile_sum[t];
                    if (w > 0) {
                        c = poisson_draw(key, t | ((uint64_t)pa.salt[p] << 48) | ((uint64_t)stage << 40), gen, PH_PLAN, mu2 * ((double)w / (double)W), s_rcp);
                        st->tile_cnt[p * stride + t] += c;
                    }
                }
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, sh);
                if (lane == 0 && c) atomicAdd(&s_acc[p], (unsigned long long)c);
            }
            __syncthreads();
            if (tid < NP && (active >> tid & 1u) && s_acc[tid]) atomicAdd((unsigned long long *)&st->plan_acc[stage][tid], s_acc[tid]);
            plan_grid_barrier(st, epoch);
            for (uint32_t p = 0; p < NP; p++) if (active >> p & 1u) K[p] += plan_ld(&st->plan_acc[stage][p]);
        }
        // what is still missing is drawn one by one; the draws of all parts form one index space over all plan threads
        uint64_t D_total = 0;
        for (uint32_t p = 0; p < NP; p++) {
            if (K[p] > s_size[p]) tree_mask |= 1u << p;
            else D_total += s_size[p] - K[p];
        }
        for (uint64_t e = (uint64_t)blockIdx.x * PLAN_THREADS + tid; e < D_total; e += (uint64_t)gridDim.x * PLAN_THREADS) {
            uint32_t p = 0;
            uint64_t j = e;
            for (;; p++) {
                const uint64_t D = K[p] > s_size[p] ? 0 : s_size[p] - K[p];
                if (j < D) break;
                j -= D;
            }
            // Uniform(0, W): 64-bit widening multiply with rejection
            Philox g(st->seed, st->compartment, gen, PH_PLAN_FILL, j | ((uint64_t)pa.salt[p] << 48));
            const uint64_t r = g.below(W);
            uint64_t lo = 0, hi = nt;  // first tile with prefix > r
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (prefix[mid] <= r) lo = mid + 1; else hi = mid;
            }
            atomicAdd((unsigned long long *)&st->tile_cnt[p * stride + lo], 1ull);
        }
    }
    if (tree_mask) plan_grid_barrier(st, epoch);  // (uniform) stage counts of the parts that fall back are about to be overwritten
    if (tree_mask && lead) {
        const uint64_t P = pow2_ceil(nt ? nt : 1);
        uint64_t *tr = st->tile_tree, *ncnt = st->tile_node;
        __syncthreads();
        for (uint64_t t = tid; t < P; t += PLAN_THREADS) tr[P + t] = t < nt ? st->tile_sum[t] : 0;
        __syncthreads();
        for (uint64_t s = P >> 1; s >= 1; s >>= 1) {
            for (uint64_t node = s + tid; node < 2 * s; node += PLAN_THREADS) tr[node] = tr[2 * node] + tr[2 * node + 1];
            __syncthreads();
        }
        for (uint32_t p = 0; p < NP; p++) {
            if (!(tree_mask >> p & 1u)) continue;
            uint64_t *cnt = st->tile_cnt + p * stride;
            if (tid == 0) ncnt[1] = s_size[p];
            __syncthreads();
            for (uint64_t s = 1; s < P; s <<= 1) {
                for (uint64_t node = s + tid; node < 2 * s; node += PLAN_THREADS) {
                    const uint64_t c = ncnt[node], wl = tr[2 * node], w = tr[node];
                    uint64_t left = 0;
                    if (c > 0 && wl > 0) {
                        if (wl == w) left = c;
                        else {
                            Philox g(st->seed, st->compartment, gen, PH_PLAN_TREE, node | ((uint64_t)pa.salt[p] << 48));
                            left = binomial(g, c, (double)wl / (double)w);
                        }
                    }
                    ncnt[2 * node] = left;
                    ncnt[2 * node + 1] = c - left;
                }
                __syncthreads();
            }
            for (uint64_t t = tid; t < nt; t += PLAN_THREADS) cnt[t] = ncnt[P + t];
            __syncthreads();
        }
    }
    plan_grid_barrier(st, epoch);  // every count is final
    if (lead && tid == 0) st->dbg[3] = vl_globaltimer();
    // ---- (3) exclusive scan of the counts: slice totals first, then every CTA scans its slice from its base
    if (NP < 4) {  // few parts: the whole CTA scans one part's slice after the other
        for (uint32_t p = 0; p < NP; p++) {
            const uint64_t *cnt = st->tile_cnt + p * stride;
            uint64_t part = 0;
            for (uint64_t t = t_lo + tid; t < t_hi; t += PLAN_THREADS) part += __ldcg(&cnt[t]);
            part = plan_block_sum(part, s_scan);
            if (tid == 0) st->plan_part[blockIdx.x][p] = part;
        }
        plan_grid_barrier(st, epoch);
        for (uint32_t p = 0; p < NP; p++) {
            uint64_t base = 0;
            for (uint32_t g = 0; g < blockIdx.x; g++) base += plan_ld(&st->plan_part[g][p]);
            plan_block_scan(st->tile_cnt + p * stride + t_lo, st->tile_base + p * stride + t_lo, t_hi - t_lo, false, base, s_scan, &s_carry);
        }
        if (lead && tid == 0) st->dbg[4] = vl_globaltimer();
        plan_leave();
        return;
    }
    // many parts: one warp per part (round robin): the slice total, then — once every CTA's totals are known — the scan of
    // the slice, 32 tiles at a time (other CTAs may have added to the counts: read past L1)
    for (uint32_t p = wid; p < NP; p += PLAN_THREADS / 32) {
        const uint64_t *cnt = st->tile_cnt + p * stride;
        uint64_t part = 0;
        for (uint64_t t = t_lo + lane; t < t_hi; t += 32) part += __ldcg(&cnt[t]);
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, sh);
        if (lane == 0) st->plan_part[blockIdx.x][p] = part;
    }
    plan_grid_barrier(st, epoch);
    for (uint32_t p = wid; p < NP; p += PLAN_THREADS / 32) {
        uint64_t run = 0;
        for (uint32_t g = 0; g < blockIdx.x; g++) run += plan_ld(&st->plan_part[g][p]);
        const uint64_t *cnt = st->tile_cnt + p * stride;
        uint64_t *base = st->tile_base + p * stride;
        for (uint64_t t0 = t_lo; t0 < t_hi; t0 += 32) {
            const uint64_t t = t0 + lane;
            const uint64_t x = t < t_hi ? __ldcg(&cnt[t]) : 0;
            uint64_t inc = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= (uint32_t)d) inc += y;
            }
            if (t < t_hi) base[t] = run + inc - x;
            run += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
    }
    if (lead && tid == 0) st->dbg[4] = vl_globaltimer();
    plan_leave();
}

// One CTA per tile of CSR positions.  The tile's records with offspring > 0 are compacted into shared memory (payload =
// haplotype id or position); a draw value r in [0, total) belongs to the record whose run [lo, hi) of the offspring CDF
// contains it.  Two ways to find it, same answer (so which one a tile takes does not change the result):
//  - rank select (totals up to RESAMPLE_RANK_MAX, every K configuration): a bit mask over the draw values has a bit at
//    every run start; the record of r is the number of set bits at or below r, i.e. one mask word, one per-word prefix
//    count and a popc.  Building it costs one shared atomicOr per record — nothing per offspring;
//  - compacted inclusive CDF + cut-point guide table on the top 11 bits of the random word (one guide lookup and ~1-2
//    CDF probes) for larger totals.
// One Philox block serves four consecutive output slots.  No global gathers.
// mode 0: write the haplotype ids of the drawn records to dst; mode 1: write the drawn infectant indices (u64).
constexpr uint32_t GUIDE_BITS = 11;
constexpr uint32_t GUIDE_SIZE = 1u << GUIDE_BITS;
constexpr uint32_t RESAMPLE_RANK_MAX = 32768;  // largest tile offspring total served by the bit mask (4 KB of shared memory)
constexpr uint32_t RANK_WORDS = RESAMPLE_RANK_MAX / 32;
static_assert(RANK_WORDS % TPB == 0, "every thread owns RANK_WORDS / TPB consecutive mask words");
// Several parts (subsamples of the same offspring map) are drawn in the same pass: the tile is staged once, every part
// has its own draw counts, output buffer and Philox stream.
struct ResampleArgs {
    uint32_t n_parts;
    uint32_t salt[MAX_PARTS];
    uint32_t *dst[MAX_PARTS];  // nullptr = the population being assembled (hap_id_next)
};
__global__ void __launch_bounds__(TPB, 4) k_resample(DevState *st, ResampleArgs ra, uint64_t *__restrict__ dst64, int mode) {
    __shared__ uint32_t s_hap[RESAMPLE_TILE];   // compacted payloads: haplotype id (mode 0) or position in the tile (mode 1)
    __shared__ uint32_t s_cdf[RESAMPLE_TILE];   // compacted inclusive CDF (guide path only)
    __shared__ uint16_t s_guide[GUIDE_SIZE];
    __shared__ uint32_t s_mask[RANK_WORDS + 1];
    __shared__ uint16_t s_pre[RANK_WORDS];      // set bits in the mask words before this one
    __shared__ uint32_t s_wbits[TPB / 32];
    __shared__ uint64_t s_pcnt[MAX_PARTS], s_pbase[MAX_PARTS];
    __shared__ uint32_t s_wsum[TPB / 32], s_wcnt[TPB / 32];
    const uint64_t nt = tile_count(st);
    const uint32_t *__restrict__ offspring = st->offspring;
    const uint2 *__restrict__ rec = st->rec;
    constexpr uint32_t PER = RESAMPLE_TILE / TPB;  // consecutive items per thread
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const PhiloxKey key = philox_key(st->seed, st->compartment);
    const uint32_t gen = st->generation;
    const uint64_t stride = st->tile_stride;
    for (uint64_t tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        // the tile's descriptors are independent loads: issue them together, decide afterwards (with several parts every
        // part's draw count and first output slot are parked in shared memory: no dependent load per part later)
        uint64_t any = 0;
        if (ra.n_parts == 1) {
            any = st->tile_cnt[tile];
        } else {
            __syncthreads();  // (a tile without draws is left without a barrier: nobody may still be reading its entries)
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
        // the range of the tile this CTA takes next: asked for now, used at the bottom of the loop to pull that tile's
        // offspring counts and records towards L2 while this tile's draws are written
        uint64_t next_start = 0;
        uint32_t next_len = 0;
        if (tile + gridDim.x < nt) tile_range(st, tile + gridDim.x, next_start, next_len);
        if (any == 0) continue;  // uniform across the CTA
        if (tsum > 0xFFFFFFFFull) {  // the tile CDF is 32-bit
            if (tid == 0) raise(st, DE_OFFSPRING_RANGE);
            continue;
        }
        if (tlen > RESAMPLE_TILE) {  // cannot happen: host tiles are only declared by a kernel that staged every tile
            if (tid == 0) raise(st, DE_TILE_OVERFLOW);
            continue;
        }
        // ---- load PER consecutive items per thread, local scans of offspring and of the non-zero flags
        uint32_t v[PER], hp[PER];
        const uint64_t p0 = start + (uint64_t)tid * PER;
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) {
            const uint64_t p = p0 + k;
            const bool in = tid * PER + k < tlen;
            v[k] = in ? offspring[p] : 0u;
            // (asked for whether or not the record has offspring: the sector comes anyway, and the load does not wait for v)
            hp[k] = !in ? 0u : mode == 0 ? rec[p].y : (uint32_t)(p - start);
        }
        uint32_t run = 0, nz = 0;
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) { run += v[k]; nz += v[k] ? 1u : 0u; }
        uint32_t inc = run, cinc = nz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d), z = __shfl_up_sync(0xFFFFFFFFu, cinc, d);
            if (lane >= (uint32_t)d) { inc += y; cinc += z; }
        }
        if (lane == 31) { s_wsum[wid] = inc; s_wcnt[wid] = cinc; }
        constexpr uint32_t WPT = RANK_WORDS / TPB;  // mask words per thread
#pragma unroll
        for (uint32_t k = 0; k < WPT; k++) s_mask[tid * WPT + k] = 0u;
        __syncthreads();
        uint32_t woff = 0, total = 0, coff = 0;
#pragma unroll
        for (int w = 0; w < TPB / 32; w++) {
            const uint32_t t = s_wsum[w], u = s_wcnt[w];
            if (w < (int)wid) { woff += t; coff += u; }
            total += t;
        }
        uint32_t cdf = woff + inc - run;   // exclusive prefix of this thread's first item
        uint32_t slot = coff + cinc - nz;  // compacted index of this thread's first non-zero item
        const bool direct = total <= RESAMPLE_RANK_MAX;
        const double scale = (double)GUIDE_SIZE / (double)total;
        if (direct) {
#pragma unroll
            for (uint32_t k = 0; k < PER; k++) {
                if (!v[k]) continue;
                atomicOr(&s_mask[cdf >> 5], 1u << (cdf & 31));
                s_hap[slot++] = hp[k];
                cdf += v[k];
            }
            __syncthreads();
            // set bits before every mask word: per-thread popc, warp scan, warp totals
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
            for (int w = 0; w < TPB / 32; w++)
                if (w < (int)wid) before += s_wbits[w];
#pragma unroll
            for (uint32_t k = 0; k < WPT; k++) { s_pre[tid * WPT + k] = (uint16_t)before; before += bits[k]; }
        } else
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) {
            if (!v[k]) continue;
            const uint32_t lo = cdf, hi = cdf + v[k];
            cdf = hi;
            s_cdf[slot] = hi;
            s_hap[slot] = hp[k];
            // guide[b] = this entry for every bucket b whose first reachable draw value (b*total) >> GUIDE_BITS lies in [lo, hi)
            uint32_t bf = (uint32_t)((double)lo * scale);
            if (bf > GUIDE_SIZE) bf = GUIDE_SIZE;
            while (bf > 0 && (uint32_t)(((uint64_t)(bf - 1) * total) >> GUIDE_BITS) >= lo) bf--;
            while (bf < GUIDE_SIZE && (uint32_t)(((uint64_t)bf * total) >> GUIDE_BITS) < lo) bf++;
            for (uint32_t b = bf; b < GUIDE_SIZE && (uint32_t)(((uint64_t)b * total) >> GUIDE_BITS) < hi; b++) s_guide[b] = (uint16_t)slot;
            slot++;
        }
        __syncthreads();
        // ---- draws of every part: output slots [base, base + c), four per Philox block.  The Philox blocks of all parts
        // form one index space dealt to the threads, so the handful of draws of a small part does not cost a pass of its own.
        uint32_t blocks_total = 0;  // (every thread computes the same prefix from shared memory: <= 16 parts)
        if (ra.n_parts > 1) {
            for (uint32_t p = 0; p < ra.n_parts; p++) {
                const uint64_t c = s_pcnt[p], b = s_pbase[p];
                blocks_total += c ? (uint32_t)(((b + c - 1) >> 2) - (b >> 2) + 1) : 0u;
            }
        } else {
            const uint64_t b = st->tile_base[tile];
            blocks_total = (uint32_t)(((b + any - 1) >> 2) - (b >> 2) + 1);
        }
        for (uint32_t e = tid; e < blocks_total; e += TPB) {
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
                uint32_t payload;
                if (direct) {
                    const uint32_t upto = s_mask[r >> 5] & (0xFFFFFFFFu >> (31 - (r & 31)));  // run starts at or below r in its word
                    payload = s_hap[(uint32_t)s_pre[r >> 5] + __popc(upto) - 1u];
                } else {
                    uint32_t i = s_guide[x >> (32 - GUIDE_BITS)];
                    while (s_cdf[i] <= r) i++;
                    payload = s_hap[i];
                }
                if (mode == 0) dst[s] = payload;
                else dst64[s] = rec[start + payload].x;
            }
        }
        if (tid * 8u < next_len) {  // 32 bytes of offspring counts and 64 bytes of records per thread
            asm volatile("prefetch.global.L2 [%0];" ::"l"(offspring + next_start + tid * 8u));
            if (mode == 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + next_start + tid * 8u));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + next_start + tid * 8u + 4u));
            }
        }
        __syncthreads();
    }
}

// Population::select (src/core/population.rs:406-430): ids of the infectants at the given indices
__global__ void __launch_bounds__(TPB) k_select(DevState *st, const uint64_t *__restrict__ idx, uint64_t m, uint32_t *__restrict__ dst) {
    const uint64_t n = st->n;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t i = idx[k];
        if (i >= n) { raise(st, DE_REPLAY); i = 0; }
        dst[k] = st->hap_id[i];
    }
}

// set_population: the assembled population becomes current (src/simulation.rs:340-345)
__global__ void k_swap_population(DevState *st) {
    uint32_t *t = st->hap_id;
    st->hap_id = st->hap_id_next;
    st->hap_id_next = t;
    st->n = st->n_next;
    st->n_sorted = 0;
}

// ------------------------------------------------------------------------------------------ garbage collection
// A record is live iff an infectant (or a detached population) references it: records hold no references to other
// records (a child shares its parent's base LIST, which is arena words, not a record).  Compaction runs only when the
// table or the arena passes its threshold, and works on a bit map of the records, one word per 32:
//   mark    one atomic OR per infectant (the map of 10^8 records is 12 MB: it stays in L2),
//   words   live records and live (merged) entries per word, two scans over the words,
//   copy    every live record is FLATTENED into the other arena (its own list, empty delta) under its new id =
//           scan value of its word + set bits below it in the word; fitness memos move along,
//   remap   every infectant looks its new id up with the same two reads.
// The cost is proportional to the live records and the population, not to the records created since the last run.
__device__ __forceinline__ uint32_t gc_new_id(const DevState *st, uint32_t id) {
    const uint32_t w = id >> 5;
    return st->gc_cbase[w] + __popc(st->gc_bits[w] & ((1u << (id & 31)) - 1u));
}
// gc_active: 0 = nothing to do, 2 = compaction: every live record is flattened into the other arena under its new id.
// Flattening the SURVIVORS every few generations is what keeps k_mutate cheap: a lineage mutates ~0.2 times per
// generation at K2, so between two compactions its delta rarely overflows and almost no record is flattened by the
// mutate kernel itself (which would copy a list per mutation of lineages that mostly die out).  (Compacting the heads
// only and leaving the lists in place was measured and dropped: 1.63 against 1.30 ms per generation at generation 1000.)
__global__ void k_gc_begin(DevState *st, int force) {
    st->gc_active = (force || st->n_hap > st->gc_hap_threshold || st->arena_top > st->gc_arena_threshold) ? 2u : 0u;
}
// keep_wildtype: record 0 exists on every handle and keeps its id (compaction); the dedup of a detached population
// (vl_pop_pack_device) never ships it
__global__ void __launch_bounds__(TPB) k_gc_clear(DevState *st, int guarded, int keep_wildtype) {
    if (guarded && !st->gc_active) return;
    const uint64_t nw = (st->n_hap + 31) >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->gc_nw = nw;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += (uint64_t)gridDim.x * blockDim.x)
        st->gc_bits[w] = (w == 0 && keep_wildtype) ? 1u : 0u;
}
__global__ void __launch_bounds__(TPB) k_gc_mark(DevState *st, const uint32_t *__restrict__ ids, uint64_t m, int use_population, int guarded, int skip_wildtype) {
    if (guarded && !st->gc_active) return;
    if (use_population) { ids = st->hap_id; m = st->n; }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t id = ids[i];
        if (skip_wildtype && id == 0) continue;
        if (id >= st->n_hap) continue;  // (the id of a pending record: nothing to keep yet)
        atomicOr(&st->gc_bits[id >> 5], 1u << (id & 31));
    }
}
// The worklist of a generation the last resample prepared (k_resample_pos<true>) refers to records too: the parents of
// the pending mutants stay live, and the population already holds the ids the pending records WILL get (record count +
// worklist position), which move down with the record count.
__global__ void __launch_bounds__(TPB) k_gc_mark_pending(DevState *st) {
    if (!st->gc_active) return;
    const uint32_t m = st->n_mut_list;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const uint32_t id = st->mut_list[i].w;
        atomicOr(&st->gc_bits[id >> 5], 1u << (id & 31));
    }
}
__global__ void __launch_bounds__(TPB) k_gc_remap_pending(DevState *st) {
    if (!st->gc_active) return;
    const uint32_t m = st->n_mut_list;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) st->mut_list[i].w = gc_new_id(st, st->mut_list[i].w);
}
__global__ void __launch_bounds__(TPB) k_gc_words(DevState *st, int guarded) {
    if (guarded && !st->gc_active) return;
    const uint64_t nw = (st->n_hap + 31) >> 5;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t bits = st->gc_bits[w], len = 0;
        st->gc_wcnt[w] = __popc(bits);
        while (bits) {
            const uint32_t b = __ffs(bits) - 1;
            bits &= bits - 1;
            len += st->head[(w << 5) + b].nd_mlen >> 8;
        }
        st->gc_wlen[w] = len;
    }
}
// One CTA per TPB bit-map words: every thread lists the live records of its word (old id, new id, new arena offset) in
// shared memory, then the CTA copies the listed records one per thread — the loads of different records overlap instead
// of being chained inside one thread's walk over its word.
constexpr uint32_t GC_LONG = 24;   // records longer than this are copied by a whole warp
constexpr uint32_t GC_LIST = 2048;  // live records a CTA pass can list (TPB words hold up to 32 * TPB; longer lists are copied by the lister)
__global__ void __launch_bounds__(TPB) k_gc_copy(DevState *st) {
    __shared__ uint32_t s_old[GC_LIST], s_new[GC_LIST];
    __shared__ unsigned long long s_off[GC_LIST];
    __shared__ uint32_t s_delta[TPB / 32][HEAD_DELTA];
    __shared__ uint32_t s_cnt;
    if (!st->gc_active) return;
    const uint64_t nw = (st->n_hap + 31) >> 5;
    auto copy_one = [&](uint64_t h, uint32_t nid, uint64_t noff) {
        const HapHead hd = st->head[h];
        uint32_t *dst = st->arena_alt + noff;
        uint32_t o = 0;
        merged_for_each(st->arena + hd.off, hd.len, hd.d, head_nd(hd), [&](uint32_t e) { dst[o++] = e; });
        HapHead nh;
        nh.off = noff;
        nh.len = o;
        nh.nd_mlen = o << 8;
        nh.raw0 = hd.raw0;
        nh.fit0 = hd.fit0;
#pragma unroll
        for (uint32_t k = 0; k < HEAD_DELTA; k++) nh.d[k] = 0xFFFFFFFFu;
        st->head_alt[nid] = nh;
        for (uint32_t p = 0; p < st->n_providers; p++) { hap_raw_alt(st, p, nid) = hap_raw(st, p, h); st->h_fit_alt[p][nid] = st->h_fit[p][h]; }
    };
    for (uint64_t w0 = (uint64_t)blockIdx.x * TPB; w0 < nw; w0 += (uint64_t)gridDim.x * TPB) {
        __syncthreads();
        if (threadIdx.x == 0) s_cnt = 0u;
        __syncthreads();
        const uint64_t w = w0 + threadIdx.x;
        uint32_t bits = w < nw ? st->gc_bits[w] : 0u;
        if (bits) {
            const uint32_t c = __popc(bits);
            const uint32_t slot = atomicAdd(&s_cnt, c);
            uint32_t nid = st->gc_cbase[w];
            uint64_t noff = st->gc_lbase[w];
            uint32_t k = 0;
            while (bits) {
                const uint32_t b = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint64_t h = (w << 5) + b;
                const uint32_t len = st->head[h].nd_mlen >> 8;
                if (slot + k < GC_LIST) { s_old[slot + k] = (uint32_t)h; s_new[slot + k] = nid; s_off[slot + k] = noff; }
                else copy_one(h, nid, noff);  // (more live records under this CTA's words than the list holds)
                k++;
                nid += 1;
                noff += len;
            }
        }
        __syncthreads();
        const uint32_t m = s_cnt < GC_LIST ? s_cnt : GC_LIST;
        // short records: one thread each; long ones are left to whole warps (nothing serial in the list's length)
        for (uint32_t k = threadIdx.x; k < m; k += TPB) {
            if ((st->head[s_old[k]].nd_mlen >> 8) <= GC_LONG) { copy_one(s_old[k], s_new[k], s_off[k]); s_old[k] = NONE32; }
        }
        __syncthreads();
        const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (uint32_t k = wid; k < m; k += TPB / 32) {
            const uint32_t h = s_old[k];
            if (h == NONE32) continue;  // (warp-uniform)
            const HapHead hd = st->head[h];
            if (lane < HEAD_DELTA) s_delta[wid][lane] = hd.d[lane];
            __syncwarp();
            const uint64_t noff = s_off[k];
            const uint32_t o = warp_flatten(st->arena + hd.off, hd.len, s_delta[wid], head_nd(hd), st->arena_alt + noff, st->L);
            __syncwarp();
            if (lane == 0) {
                const uint32_t nid = s_new[k];
                HapHead nh;
                nh.off = noff;
                nh.len = o;
                nh.nd_mlen = o << 8;
                nh.raw0 = hd.raw0;
                nh.fit0 = hd.fit0;
#pragma unroll
                for (uint32_t q = 0; q < HEAD_DELTA; q++) nh.d[q] = 0xFFFFFFFFu;
                st->head_alt[nid] = nh;
                for (uint32_t p = 0; p < st->n_providers; p++) { hap_raw_alt(st, p, nid) = hap_raw(st, p, h); st->h_fit_alt[p][nid] = st->h_fit[p][h]; }
            }
        }
    }
}
__global__ void __launch_bounds__(TPB) k_gc_remap(DevState *st, uint32_t *ids, uint64_t m, int use_population, const uint64_t *gc_totals) {
    if (!st->gc_active) return;
    if (use_population) { ids = st->hap_id; m = st->n; }
    const uint64_t n_old = st->n_hap, n_live = gc_totals[0];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t id = ids[i];
        // (an id at or beyond the record count is the id a pending record will get: see k_gc_mark_pending)
        ids[i] = id < n_old ? gc_new_id(st, id) : (uint32_t)(n_live + (id - n_old));
    }
}
// gc_totals[0] = live records, gc_totals[1] = live arena words (written by the two scans)
__global__ void k_gc_end(DevState *st, const uint64_t *gc_totals) {
    if (!st->gc_active) return;
    st->n_hap = gc_totals[0];
    st->arena_top = gc_totals[1];
#define VL_SWAP(a, b) { auto t = st->a; st->a = st->b; st->b = t; }
    VL_SWAP(head, head_alt) VL_SWAP(arena, arena_alt)
    for (uint32_t p = 0; p < st->n_providers; p++) { VL_SWAP(h_raw[p], h_raw_alt[p]) VL_SWAP(h_fit[p], h_fit_alt[p]) }
#undef VL_SWAP
    st->gc_runs += 1;
    st->gc_active = 0;
}

// ------------------------------------------------------------------------------------------ small helpers
// clears n 32-bit words.  Used instead of cudaMemsetAsync on the enqueue paths: a device memset is one of the commands
// that can implicitly serialise streams of one process, and compartments of one process must be able to run side by
// side (one waits on the device for the other's migrants).
__global__ void __launch_bounds__(TPB) k_zero_words(uint32_t *p, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = 0u;
}
__global__ void __launch_bounds__(TPB) k_fill_u32(uint32_t *p, uint64_t n, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
// offspring map from the host (u64 per infectant, identity layout) -> device u32 per position
__global__ void __launch_bounds__(TPB) k_u64_to_u32(DevState *st, const uint64_t *__restrict__ src, uint32_t *__restrict__ dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = src[i];
        if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0; }
        dst[i] = (uint32_t)v;
    }
}
// offspring per CSR position -> Vec<usize> per infectant (dst pre-zeroed: uninfected infectants have 0 offspring)
__global__ void __launch_bounds__(TPB) k_unsort_offspring(DevState *st, uint64_t *__restrict__ dst) {
    const uint64_t m = st->n_sorted;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (uint64_t)gridDim.x * blockDim.x) dst[st->rec[p].x] = st->offspring[p];
}
__global__ void __launch_bounds__(TPB) k_unsort_offspring32(DevState *st, uint32_t *__restrict__ dst) {
    const uint64_t m = st->n_sorted;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (uint64_t)gridDim.x * blockDim.x) dst[st->rec[p].x] = st->offspring[p];
}
__global__ void __launch_bounds__(TPB) k_extract_hosts(DevState *st, uint32_t *__restrict__ dst) {
    const uint64_t m = st->n_sorted;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (uint64_t)gridDim.x * blockDim.x) dst[p] = st->rec[p].x;
}
// per-infectant replication terms for inspection: mapped fitness and host sum as BasicHost::replicate forms them,
// lambda as the replicate kernels left it at the record's CSR position (0.0 where Poisson::new fails); outputs are
// pre-filled with NaN for uninfected infectants; ordered slices required
__global__ void __launch_bounds__(TPB) k_replication_terms(DevState *st, double *fit_out, double *hs_out, double *lam_out) {
    const uint64_t H = st->H;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < H; h += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = st->offsets[h], e = st->offsets[h + 1];
        if (e == b) continue;
        const int s = find_spec(st, h);
        double S = 0.0;
        for (uint32_t a = b; a < e; a++) S += reproductive_fitness(st, s, st->rec[a].y);
        for (uint32_t a = b; a < e; a++) {
            const double f = reproductive_fitness(st, s, st->rec[a].y);
            const uint32_t i = st->rec[a].x;
            fit_out[i] = f;
            hs_out[i] = S;
            lam_out[i] = st->lam[a];
        }
    }
}
__global__ void __launch_bounds__(TPB) k_gather_raw(DevState *st, uint32_t provider, double *out) {
    const uint64_t n = st->n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = hap_raw(st, provider, st->hap_id[i]);
}

// sampler test hooks (distribution tests of the device samplers the kernels above call)
__global__ void k_test_poisson(uint64_t seed, double lambda, uint64_t n, uint64_t *out) {
    __shared__ double s_rcp[POISSON_RCP];
    poisson_rcp_init(s_rcp);
    __syncthreads();
    const PhiloxKey key = philox_key(seed, 0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = poisson_draw(key, i, 0, PH_TEST, lambda, s_rcp);
}
// out[i] = the f32-guarded inversion, out[n + i] = the plain f64 inversion of the same uniform: must be equal draw for draw
__global__ void k_test_poisson_fast(uint64_t seed, double lambda, uint64_t n, uint64_t *out) {
    __shared__ double s_rcp[POISSON_RCP];
    poisson_rcp_init(s_rcp);
    __syncthreads();
    const PhiloxKey key = philox_key(seed, 0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        // lambda <= 0 selects a per-draw mean spread over (0, 12) so that the guard band is exercised at every scale
        const double l = lambda > 0.0 ? lambda : 12.0 * u01_open_from(philox_block(key, i, 1, PH_TEST, 0).x, philox_block(key, i, 1, PH_TEST, 0).y);
        out[i] = poisson_draw_small_fast(key, i, 0, PH_TEST, l, s_rcp, poisson_kmax(lambda > 0.0 ? lambda : 12.0));
        const uint4 b = philox_block(key, i, 0, PH_TEST, 0);
        out[n + i] = poisson_inversion_f64(u01_from(b.x, b.y), l, s_rcp);
    }
}
__global__ void k_test_binomial(uint64_t seed, uint64_t nn, double p, uint64_t n, uint64_t *out, BinvConst c) {
    const PhiloxKey key = philox_key(seed, 0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (c.ok) {
            const uint4 b0 = philox_block(key, i, 0, PH_TEST, 0);
            // the product path's integer thresholds against the f64 search they were derived from; a disagreement
            // is reported as a value no binomial can take
            const uint64_t U = (((uint64_t)b0.w << 32) | b0.z) >> 11;
            uint32_t k = 0;
            while (k < (uint32_t)BINV_THRESHOLDS && U >= c.thr[k]) k++;
            const uint32_t ref = binv_draw(c, u01_from(b0.z, b0.w), key, i, 0, PH_TEST);
            out[i] = (k < (uint32_t)BINV_THRESHOLDS && k != ref) ? (1ull << 62) : ref;
        } else {
            Philox g(seed, 0, 0, PH_TEST, i);
            out[i] = binomial(g, nn, p);
        }
    }
}
__global__ void k_test_below(uint64_t seed, uint64_t bound, uint64_t n, uint64_t *out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Philox g(seed, 0, 0, PH_TEST, i);
        out[i] = g.below(bound);
    }
}

}  // namespace vl
