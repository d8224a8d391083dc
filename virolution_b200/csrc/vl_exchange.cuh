// vl_exchange.cuh — the transmission block fused with its transport: every origin writes its migrants straight into the
// targets' device memory over NVLink (peer pointers from CUDA IPC), with sizes that never leave the device.
//
// Reference: src/runner.rs:401-416 — new[target] = concat over origin of subsample(sim[origin], offspring[origin], T[target][origin]).
// One compartment per GPU/process.  Every rank owns a receive buffer (two parities x one block per origin); a block is
//     header {m migrants, w entry words} | len[m_cap] | raw[m_cap * P] | entries[w_cap]
// in the format of the migrant images (vl_migrate.cuh).  Per generation a rank
//   1. draws its `world` parts in one plan + resample pass (device-resident sizes),
//   2. k_ex_lens / scan / k_ex_copy: writes length, raw fitness and entries of every off-diagonal migrant directly into the
//      target's block (peer stores), then k_ex_publish writes the header and, after a system fence, the block's flag,
//   3. k_ex_wait: waits on the flags of its own blocks (a device-side wait: the host is not involved),
//   4. k_ex_in_*: turns the received blocks into haplotype records and the new population, origin after origin.
// No collective call, no host synchronisation, no size ever read by the host.  A generation's blocks use parity
// (seq & 1): an origin cannot start the push of generation g + 2 before its target has consumed generation g, because
// that push comes after the origin's import of g + 1, which needs the target's push of g + 1, which the target enqueues
// after its import of g.
#pragma once
#include "vl_migrate.cuh"
#include "vl_fused.cuh"

namespace vl {

constexpr uint32_t EX_MAX_WORLD = MAX_PARTS;
constexpr uint64_t EX_HEADER_BYTES = 64;
constexpr uint32_t MIG_WILD_META = 0xFFFFFFFFu;

// where one (target <- origin) block lives, for both parities; pointers are valid on the GPU that dereferences them
struct ExBlock {
    uint64_t *header[2];  // {m, w, m announced early (k_ex_sizes)}
    uint32_t *len[2];
    uint4 *meta[2];       // fast path, per migrant: {first word inside entries[] (MIG_WILD_META for the wildtype), base list length,
                          // HapHead::nd_mlen, 0}; its words are the base list followed by the nd delta entries of its head,
                          // and the words of a block are not in migrant order
    unsigned long long *eflag[2];  // raised by the origin as soon as its part sizes are decided
    double *raw[2];
    uint32_t *entries[2];
    unsigned long long *flag[2];  // written last by the origin: the sequence number of the generation the block holds
    uint64_t m_cap, w_cap;
};
struct ExArgs {
    uint32_t world, rank, parity, n_providers;
    uint64_t seq;
    // sharded compartment: every rank's offspring total of this generation, written by its owner into every buffer
    uint64_t *tot_out[EX_MAX_WORLD][2];            // my slot in every rank's buffer (own buffer included)
    unsigned long long *tot_flag_out[EX_MAX_WORLD][2];
    uint64_t *tot_in[2];                           // world slots in my buffer
    unsigned long long *tot_flag_in[2];
    uint64_t shard_seed;                           // the same on every rank: the split of the draws must be the same everywhere
    ExBlock out[EX_MAX_WORLD];        // my block on every target (peer memory), [rank] unused
    ExBlock in[EX_MAX_WORLD];         // the block of every origin in my own buffer, [rank] unused
    const uint32_t *part[EX_MAX_WORLD];  // ids of the part drawn for every target (my memory)
    const double *own_fit;               // mapped fitness of the infectants of my own part (device-resident generation), or nullptr
};
// device scratch of one exchange (my memory)
struct ExScratch {
    uint64_t part_start[EX_MAX_WORLD + 1];  // concatenated off-diagonal migrants: first index of every target's part
    uint64_t n_out;                         // = part_start[world]
    uint64_t in_start[EX_MAX_WORLD + 1];    // new population: first slot of every origin's part (own part included)
    uint64_t imp_start[EX_MAX_WORLD + 1];   // concatenated imported migrants (own part excluded)
    uint64_t n_imp, w_imp;
    uint64_t totals[2];                     // records / words needed by the import (from the scans)
    uint64_t reserve[2];
    unsigned long long wcur[EX_MAX_WORLD];  // fast path: entry words written so far into every target's block
    uint64_t own_off;                       // fast path: first slot of my own part in the new population (= in_start[rank])
    unsigned int exp_done;                  // fast path: CTAs of k_ex_export that have left (the last one publishes the blocks)
};

__device__ __forceinline__ uint32_t ex_find(const uint64_t *start, uint32_t world, uint64_t idx) {
    uint32_t t = 0;
    while (t + 1 < world && idx >= start[t + 1]) t++;
    return t;
}

// concatenation order of what leaves (targets ascending, own part skipped)
__global__ void k_ex_out_starts(DevState *st, ExArgs ex, ExScratch *sc) {
    uint64_t acc = 0;
    for (uint32_t t = 0; t < ex.world; t++) {
        sc->part_start[t] = acc;
        if (t != ex.rank) {
            uint64_t m = st->plan_size[t];
            if (m > ex.out[t].m_cap) { raise(st, DE_INF_CAPACITY); m = ex.out[t].m_cap; }
            acc += m;
        }
    }
    sc->part_start[ex.world] = acc;
    sc->n_out = acc;
}
// length and raw fitness of every leaving migrant go straight to the target; the lengths also feed the word scan
__global__ void __launch_bounds__(TPB) k_ex_lens(DevState *st, ExArgs ex, const ExScratch *sc, uint32_t *__restrict__ words) {
    const uint64_t M = sc->n_out;
    const uint32_t P = ex.n_providers, par = ex.parity;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M; idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t t = ex_find(sc->part_start, ex.world, idx);
        const uint64_t k = idx - sc->part_start[t];
        const uint32_t id = ex.part[t][k];
        // the record's one-line view carries its length and the raw fitness of provider 0: one random line per migrant
        const HapHead hd = st->head[id];
        const uint32_t len = id == 0 ? 0u : head_mlen(hd);  // entries of the flattened record
        ex.out[t].len[par][k] = id == 0 ? MIG_WILDTYPE : len;
        for (uint32_t p = 0; p < P; p++) ex.out[t].raw[par][k * P + p] = p == 0 ? hd.raw0 : hap_raw(st, p, id);
        words[idx] = len;
    }
}
// entries of every leaving migrant, first gathered into a local staging buffer in concatenation order (random reads of
// the arena, local writes), then streamed to the targets word by word: consecutive threads write consecutive words of
// the same block, so what crosses NVLink are full 128-byte lines instead of one short store per migrant
__global__ void __launch_bounds__(TPB) k_ex_copy(DevState *st, ExArgs ex, const ExScratch *sc, const uint64_t *__restrict__ woff,
                                                 uint32_t *__restrict__ stage, uint64_t stage_cap) {
    const uint64_t M = sc->n_out;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M; idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t t = ex_find(sc->part_start, ex.world, idx);
        const uint32_t id = ex.part[t][idx - sc->part_start[t]];
        if (id == 0) continue;
        const HapHead hd = st->head[id];  // (the line k_ex_lens just read: an L2 hit)
        const uint32_t len = head_mlen(hd);
        if (woff[idx] + len > stage_cap) { raise(st, DE_ARENA_CAPACITY); continue; }  // (the target's block would overflow too)
        flatten_record(st, hd, stage + woff[idx]);
    }
}
__global__ void __launch_bounds__(TPB) k_ex_stream(DevState *st, ExArgs ex, const ExScratch *sc, const uint64_t *__restrict__ woff,
                                                   const uint32_t *__restrict__ stage, uint64_t stage_cap) {
    __shared__ uint64_t s_wstart[EX_MAX_WORLD + 1];
    if (threadIdx.x <= ex.world) s_wstart[threadIdx.x] = woff[sc->part_start[threadIdx.x]];
    __syncthreads();
    const uint64_t W = s_wstart[ex.world] < stage_cap ? s_wstart[ex.world] : stage_cap;
    const uint32_t par = ex.parity;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < W; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t t = ex_find(s_wstart, ex.world, g);
        const uint64_t w0 = g - s_wstart[t];
        if (w0 >= ex.out[t].w_cap) { if (w0 == ex.out[t].w_cap) raise(st, DE_ARENA_CAPACITY); continue; }  // block too small: words_per_migrant
        ex.out[t].entries[par][w0] = stage[g];
    }
}
// headers, then (after a system-scope fence: the stores of the kernels before this one are complete, this makes them
// visible before the flag) the flags the targets wait on
__global__ void k_ex_publish(DevState *st, ExArgs ex, const ExScratch *sc, const uint64_t *__restrict__ woff) {
    const uint32_t t = threadIdx.x;
    if (t >= ex.world || t == ex.rank) return;
    const uint64_t a = sc->part_start[t], b = sc->part_start[t + 1];
    uint64_t w = woff[b] - woff[a];  // exclusive scan with the total written at [n_out]
    if (w > ex.out[t].w_cap) w = 0;  // (already reported by k_ex_copy; keep the receiver inside its block)
    ex.out[t].header[ex.parity][0] = b - a;
    ex.out[t].header[ex.parity][1] = w;
    __threadfence_system();
    *(volatile unsigned long long *)ex.out[t].flag[ex.parity] = ex.seq;
}
// device-side wait for every origin's block of this generation (bounded: a lost peer becomes an error, not a hang)
__global__ void k_ex_wait(DevState *st, ExArgs ex, unsigned long long timeout_ns) {
    const uint32_t o = threadIdx.x;
    if (o >= ex.world || o == ex.rank) return;
    const unsigned long long t0 = vl_globaltimer();
    volatile unsigned long long *f = (volatile unsigned long long *)ex.in[o].flag[ex.parity];
    while (*f < ex.seq) {
        if (vl_globaltimer() - t0 > timeout_ns) { raise(st, DE_EXCHANGE_TIMEOUT); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}
// layout of the new population (origins ascending, own part at its rank) and of the concatenated imported migrants
__global__ void k_ex_in_starts(DevState *st, ExArgs ex, ExScratch *sc) {
    uint64_t pop = 0, imp = 0, w = 0;
    for (uint32_t o = 0; o < ex.world; o++) {
        sc->in_start[o] = pop;
        sc->imp_start[o] = imp;
        if (o == ex.rank) {
            pop += st->plan_size[o];
        } else {
            uint64_t m = ((volatile uint64_t *)ex.in[o].header[ex.parity])[0], wo = ((volatile uint64_t *)ex.in[o].header[ex.parity])[1];
            if (m > ex.in[o].m_cap || wo > ex.in[o].w_cap) { raise(st, DE_REPLAY); m = 0; wo = 0; }
            pop += m;
            imp += m;
            w += wo;
        }
    }
    sc->in_start[ex.world] = pop;
    sc->imp_start[ex.world] = imp;
    sc->n_imp = imp;
    sc->w_imp = w;
    if (pop > st->cap_inf) { raise(st, DE_INF_CAPACITY); pop = st->cap_inf; }
    st->n_next = pop;
}
__global__ void __launch_bounds__(TPB) k_ex_in_words(ExArgs ex, const ExScratch *sc, uint32_t *__restrict__ words, uint32_t *__restrict__ recs) {
    const uint64_t M = sc->n_imp;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M; idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t o = ex_find(sc->imp_start, ex.world, idx);
        const uint32_t l = __ldcg(ex.in[o].len[ex.parity] + (idx - sc->imp_start[o]));
        words[idx] = l == MIG_WILDTYPE ? 0u : l;
        recs[idx] = l == MIG_WILDTYPE ? 0u : 1u;
    }
}
__global__ void k_ex_reserve(DevState *st, ExScratch *sc) {
    const uint64_t id0 = st->n_hap, a0 = st->arena_top;
    sc->reserve[0] = NONE32;
    if (sc->totals[1] != sc->w_imp) { raise(st, DE_REPLAY); return; }
    if (id0 + sc->totals[0] > st->cap_hap) { raise(st, DE_HAP_CAPACITY); return; }
    if (a0 + sc->totals[1] > st->cap_arena) { raise(st, DE_ARENA_CAPACITY); return; }
    st->n_hap = id0 + sc->totals[0];
    st->arena_top = a0 + sc->totals[1];
    sc->reserve[0] = id0;
    sc->reserve[1] = a0;
}
// imported migrants become records of their own (like every mutant does) and members of the new population
__global__ void __launch_bounds__(TPB) k_ex_import(DevState *st, ExArgs ex, const ExScratch *sc, const uint32_t *__restrict__ rec_off,
                                                   const uint64_t *__restrict__ word_off) {
    const uint64_t M = sc->n_imp;
    const uint64_t id0 = sc->reserve[0], a0 = sc->reserve[1];
    const bool ok = id0 != NONE32;
    const uint32_t P = ex.n_providers, par = ex.parity;
    const uint64_t cap = st->cap_inf;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < M; idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t o = ex_find(sc->imp_start, ex.world, idx);
        const uint64_t k = idx - sc->imp_start[o];
        const uint64_t slot = sc->in_start[o] + k;
        const uint32_t l = __ldcg(ex.in[o].len[par] + k);
        uint32_t id = 0;
        if (l != MIG_WILDTYPE && ok) {
            id = (uint32_t)(id0 + rec_off[idx]);
            const uint64_t off = a0 + word_off[idx];
            for (uint32_t p = 0; p < P; p++) set_fitness(st, p, id, __ldcg(ex.in[o].raw[par] + k * P + p));
            const uint32_t *src = ex.in[o].entries[par] + (word_off[idx] - word_off[sc->imp_start[o]]);
            uint32_t *dst = st->arena + off;
            for (uint32_t q = 0; q < l; q++) dst[q] = __ldcg(src + q);
            write_flat_head(st, id, off, l);
        }
        if (slot < cap) {
            st->hap_id_next[slot] = id;
            // the mapped reproductive fitness travels with the infectant from here on (vl_fused.cuh)
            const int pi0 = st->specs[0].repro;
            st->pop_fit_next[slot] = pi0 >= 0 ? st->h_fit[pi0][id] : 1.0;
        }
    }
}
__global__ void __launch_bounds__(TPB) k_ex_own(DevState *st, ExArgs ex, const ExScratch *sc) {
    const uint64_t m = st->plan_size[ex.rank], s0 = sc->in_start[ex.rank], cap = st->cap_inf;
    const uint32_t *__restrict__ src = ex.part[ex.rank];
    const int pi0 = st->specs[0].repro;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x)
        if (s0 + k < cap) {
            st->hap_id_next[s0 + k] = src[k];
            st->pop_fit_next[s0 + k] = ex.own_fit ? ex.own_fit[k] : (pi0 >= 0 ? st->h_fit[pi0][src[k]] : 1.0);
        }
}

// ---- fast path: compartments that run the device-resident generation (vl_fused.cuh) ---------------------------------------
// Three kernels between the plan and the new population (the general kernels above: fourteen), and no pass that only
// computes sizes:
//   k_ex_sizes         the part sizes are known as soon as the plan has run: every target is told at once (header word + early
//                      flag), then the same warp learns how many migrants each origin will send -> the layout of the new
//                      population.  My OWN part can then be drawn straight into place (k_resample_pos, which also prepares the
//                      children's next infection: no k_ex_own copy, no k_infect_sum pass next generation)
//   k_ex_export        every leaving migrant in ONE pass (see the kernel); the last CTA to leave writes the block headers and,
//                      after a system-scope fence, the flags
//   k_ex_import_fused  waits for the origins' flags (bounded), then records, arena words and population slots of the
//                      imported migrants in ONE pass (ids and words reserved per CTA), their carried fitness, and their
//                      infection for the next generation
__device__ __forceinline__ void ex_sizes_body(DevState *st, const ExArgs &ex, ExScratch *sc) {
    __shared__ uint64_t s_m[EX_MAX_WORLD];
    const uint32_t t = threadIdx.x;
    if (t < ex.world) {
        uint64_t m = t == ex.rank ? 0 : st->plan_size[t];
        if (t != ex.rank && m > ex.out[t].m_cap) { raise(st, DE_INF_CAPACITY); m = ex.out[t].m_cap; }
        s_m[t] = m;
        sc->wcur[t] = 0ull;
        if (t != ex.rank) {
            ((volatile uint64_t *)ex.out[t].header[ex.parity])[2] = m;
            __threadfence_system();
            *(volatile unsigned long long *)ex.out[t].eflag[ex.parity] = ex.seq;
        }
    }
    __syncthreads();
    if (t == 0) {  // concatenation order of what leaves (targets ascending, own part skipped)
        uint64_t acc = 0;
        for (uint32_t q = 0; q < ex.world; q++) { sc->part_start[q] = acc; acc += s_m[q]; }
        sc->part_start[ex.world] = acc;
        sc->n_out = acc;
        sc->exp_done = 0u;
    }
}
__device__ __forceinline__ void ex_wait_sizes_body(DevState *st, const ExArgs &ex, ExScratch *sc, unsigned long long timeout_ns) {
    __shared__ uint64_t s_m[EX_MAX_WORLD];
    const uint32_t o = threadIdx.x;
    if (o < ex.world) {
        if (o == ex.rank) {
            s_m[o] = st->plan_size[o];
        } else {
            const unsigned long long t0 = vl_globaltimer();
            volatile unsigned long long *f = (volatile unsigned long long *)ex.in[o].eflag[ex.parity];
            while (*f < ex.seq) {
                if (vl_globaltimer() - t0 > timeout_ns) { raise(st, DE_EXCHANGE_TIMEOUT); break; }
                __nanosleep(100);
            }
            __threadfence_system();
            uint64_t m = ((volatile uint64_t *)ex.in[o].header[ex.parity])[2];
            if (m > ex.in[o].m_cap) { raise(st, DE_REPLAY); m = 0; }
            s_m[o] = m;
        }
    }
    __syncthreads();
    if (o != 0) return;
    uint64_t pop = 0, imp = 0;  // layout of the new population (origins ascending, own part at its rank) and of the imported migrants
    for (uint32_t q = 0; q < ex.world; q++) {
        sc->in_start[q] = pop;
        sc->imp_start[q] = imp;
        pop += s_m[q];
        if (q != ex.rank) imp += s_m[q];
    }
    sc->in_start[ex.world] = pop;
    sc->imp_start[ex.world] = imp;
    sc->n_imp = imp;
    sc->own_off = sc->in_start[ex.rank];
    if (pop > st->cap_inf) { raise(st, DE_INF_CAPACITY); pop = st->cap_inf; }
    st->n_next = pop;
}
// (one launch, one warp: announce my part sizes, then learn every origin's)
__global__ void k_ex_sizes(DevState *st, ExArgs ex, ExScratch *sc, unsigned long long timeout_ns) {
    ex_sizes_body(st, ex, sc);
    __syncthreads();
    ex_wait_sizes_body(st, ex, sc, timeout_ns);
}
constexpr int EXP_TPB = 256;
// last record of the round whose first word is at or before word g (s_pref: exclusive prefix sums of the round's lengths)
__device__ __forceinline__ uint32_t ex_owner(const uint32_t *s_pref, uint32_t g) {
    uint32_t lo = 0, hi = EXP_TPB;  // s_pref[lo] <= g < s_pref[hi] (s_pref[EXP_TPB] = total)
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_pref[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}
// A record travels as it is: its (possibly shared) base list and the delta entries of its head, 16 bytes of lengths and
// its raw fitness values; the receiver rebuilds the head.  Shared-memory staging of a round (dynamic): the words in
// destination order.  What a round sends to one target is one contiguous, 16-byte aligned run of words: it leaves with
// ONE bulk copy (cp.async.bulk shared -> global, the peer's memory being the global side), or, bulk == 0, with 16-byte
// stores of consecutive threads.
constexpr uint32_t EXP_STAGE = 8192;  // words a round can stage (else the round's words are stored one by one)
constexpr uint32_t EXP_SMEM_BYTES = EXP_STAGE * 4;
__device__ __forceinline__ void bulk_store(void *dst_global, const void *src_shared, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst_global)),
                 "r"((uint32_t)__cvta_generic_to_shared(src_shared)), "r"(bytes)
                 : "memory");
}
__global__ void __launch_bounds__(EXP_TPB) k_ex_export(DevState *st, ExArgs ex, ExScratch *sc, int bulk) {
    extern __shared__ __align__(128) uint8_t s_exp_dyn[];
    uint32_t *s_stage = (uint32_t *)s_exp_dyn;
    __shared__ uint32_t s_pref[EXP_TPB + 1], s_pos[EXP_TPB], s_blen[EXP_TPB];
    __shared__ uint32_t s_delta[EXP_TPB][HEAD_DELTA];
    __shared__ uint64_t s_src[EXP_TPB];
    __shared__ uint32_t s_sum[EX_MAX_WORLD], s_first[EX_MAX_WORLD], s_soff[EX_MAX_WORLD + 1];
    __shared__ unsigned long long s_base[EX_MAX_WORLD];
    __shared__ uint32_t s_w[EXP_TPB / 32];
    const uint64_t M = sc->n_out;
    const uint32_t P = ex.n_providers, par = ex.parity, W = ex.world;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t *__restrict__ arena = st->arena;
    for (uint64_t c0 = (uint64_t)blockIdx.x * EXP_TPB; c0 < M; c0 += (uint64_t)gridDim.x * EXP_TPB) {
        const uint64_t idx = c0 + tid;
        const bool active = idx < M;
        uint32_t t = 0, id = 0, len = 0, blen = 0, nd = 0;
        uint64_t k = 0;
        HapHead hd;
        hd.off = 0; hd.len = 0u; hd.nd_mlen = 0u; hd.raw0 = 1.0;
        if (active) {
            t = ex_find(sc->part_start, W, idx);
            k = idx - sc->part_start[t];
            id = ex.part[t][k];
            hd = st->head[id];  // ONE 64-byte line
            if (id != 0) { blen = hd.len; nd = head_nd(hd); len = blen + nd; }
        }
        uint32_t inc = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        if (bulk && tid < W) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // (my copies of the previous round have read their data)
        __syncthreads();  // (the shared arrays of the previous round are free)
        if (lane == 31) s_w[wid] = inc;
        if (tid < W) { s_sum[tid] = 0u; s_first[tid] = 0u; }
        __syncthreads();
        uint32_t pre = inc - len, tot = 0;
#pragma unroll
        for (int w = 0; w < EXP_TPB / 32; w++) {
            const uint32_t v = s_w[w];
            if (w < (int)wid) pre += v;
            tot += v;
        }
        if (active) {
            if (len) atomicAdd(&s_sum[t], len);
            if (tid == 0 || idx == sc->part_start[t]) s_first[t] = pre;  // the first migrant of target t in this round
        }
        __syncthreads();
        // every target's words of the round: a multiple of four words reserved in its block (so that every run starts
        // 16-byte aligned there) and a 16-byte aligned run of the staging area
        if (tid < W && s_sum[tid]) s_base[tid] = atomicAdd(&sc->wcur[tid], (unsigned long long)((s_sum[tid] + 3u) & ~3u));
        if (tid == 0) {
            uint32_t acc = 0;
            for (uint32_t q = 0; q < W; q++) { s_soff[q] = acc; acc += (s_sum[q] + 3u) & ~3u; }
            s_soff[W] = acc;
        }
        __syncthreads();
        const bool staged = s_soff[W] <= EXP_STAGE;
        s_pref[tid] = pre;
        if (tid == 0) s_pref[EXP_TPB] = tot;
        if (active) {
            const uint32_t in_run = pre - s_first[t];
            unsigned long long woff = len ? s_base[t] + in_run : 0ull;
            bool fits = true;
            if (woff + len + 3 > ex.out[t].w_cap) { raise(st, DE_ARENA_CAPACITY); fits = false; woff = 0ull; blen = 0u; }  // block too small: words_per_migrant
            s_src[tid] = hd.off;
            s_blen[tid] = blen;
            s_pos[tid] = staged ? s_soff[t] + in_run : (fits ? (uint32_t)woff : 0xFFFFFFFFu);
#pragma unroll
            for (uint32_t q = 0; q < HEAD_DELTA; q++) s_delta[tid][q] = hd.d[q];
            if (P) ex.out[t].raw[par][k * P] = hd.raw0;
            for (uint32_t p = 1; p < P; p++) ex.out[t].raw[par][k * P + p] = hap_raw(st, p, id);
            ex.out[t].meta[par][k] = make_uint4(id == 0 ? MIG_WILD_META : (uint32_t)woff, blen, fits ? hd.nd_mlen : 0u, 0u);
        }
        __syncthreads();
        // the round's words over the whole CTA (four per thread in flight): base lists from the arena, deltas from the heads
        for (uint32_t g0 = tid; g0 < tot; g0 += 4 * EXP_TPB) {
            uint32_t v[4], j4[4], r4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t g = g0 + u * EXP_TPB;
                if (g < tot) {
                    const uint32_t j = ex_owner(s_pref, g), r = g - s_pref[j], bl = s_blen[j];
                    j4[u] = j;
                    r4[u] = r;
                    v[u] = r < bl ? arena[s_src[j] + r] : s_delta[j][(r - bl) & (HEAD_DELTA - 1)];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (g0 + u * EXP_TPB >= tot) continue;
                if (staged) s_stage[s_pos[j4[u]] + r4[u]] = v[u];
                else if (s_pos[j4[u]] != 0xFFFFFFFFu) {
                    const uint32_t q = ex_find(sc->part_start, W, c0 + j4[u]);
                    ex.out[q].entries[par][s_pos[j4[u]] + r4[u]] = v[u];
                }
            }
        }
        if (bulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (my generic-proxy writes, before the barrier the copying threads pass)
        __syncthreads();
        if (!staged) continue;
        if (bulk) {
            if (tid < W) {
                const uint32_t words = s_soff[tid + 1] - s_soff[tid];
                if (words && s_base[tid] + words <= ex.out[tid].w_cap) {
                    bulk_store(ex.out[tid].entries[par] + s_base[tid], s_stage + s_soff[tid], words * 4u);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else {
            for (uint32_t q = 0; q < W; q++) {
                const uint32_t nq = (s_soff[q + 1] - s_soff[q]) / 4;
                if (nq && s_base[q] + nq * 4ull <= ex.out[q].w_cap) {
                    const uint4 *src = (const uint4 *)(s_stage + s_soff[q]);
                    uint4 *dst = (uint4 *)(ex.out[q].entries[par] + s_base[q]);
                    for (uint32_t c = tid; c < nq; c += EXP_TPB) dst[c] = src[c];
                }
            }
        }
    }
    // the last CTA to leave publishes the blocks: headers, then (after a system-scope fence) the flags the targets wait on
    __shared__ uint32_t s_last;
    if (bulk && tid < W) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // (all my copies are complete ...)
        asm volatile("fence.proxy.async;" ::: "memory");           // (... and ordered before what this thread writes next)
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        s_last = atomicAdd(&sc->exp_done, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (tid < W && tid != ex.rank) {
        uint64_t w = *(volatile unsigned long long *)&sc->wcur[tid];
        if (w > ex.out[tid].w_cap) w = ex.out[tid].w_cap;
        ex.out[tid].header[par][0] = sc->part_start[tid + 1] - sc->part_start[tid];
        ex.out[tid].header[par][1] = w;
        __threadfence_system();
        *(volatile unsigned long long *)ex.out[tid].flag[par] = ex.seq;
    }
}
// the worklist items staged so far (the children of my own part) and the ids of the records k_mutate (range 1) gives them
__global__ void k_ex_own_items(DevState *st) {
    const uint32_t n = st->n_mut_list;
    st->mut_own = n;
    st->mut_own_base = atomicAdd(&st->n_hap, (unsigned long long)n);
}
// imported migrants become records of their own and members of the new population; prefill: they are infected for the
// next generation here (their slot in the new population is known), like the children k_resample_pos draws
__global__ void __launch_bounds__(EXP_TPB, 4) k_ex_import_fused(DevState *st, ExArgs ex, const ExScratch *sc, int prefill, unsigned long long timeout_ns) {
    __shared__ uint4 s_items[EXP_TPB];
    __shared__ uint32_t s_clcnt[2];
    __shared__ uint32_t s_pref[EXP_TPB + 1], s_src[EXP_TPB], s_org[EXP_TPB];
    __shared__ uint32_t s_wr[EXP_TPB / 32], s_ww[EXP_TPB / 32];
    __shared__ unsigned long long s_rbase, s_wbase;
    const uint64_t M = sc->n_imp;
    const uint32_t P = ex.n_providers, par = ex.parity;
    const uint64_t cap = st->cap_inf;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int pi0 = st->specs[0].repro;
    InfectCtx ic;
    ic.H32 = (uint32_t)st->H;
    ic.thr = ic.H32 ? (uint32_t)(-ic.H32) % ic.H32 : 0u;
    ic.gen = st->generation + 1;
    ic.key = philox_key(st->seed, st->compartment);
    ic.host_id = st->host_id;
    ic.hsum = st->hsum[st->hsum_parity ^ 1u];
    ic.keep = l2_keep_policy();
    ic.stream = l2_stream_policy();
    if (tid == 0) s_clcnt[0] = 0u;
    if (timeout_ns && tid < ex.world && tid != ex.rank) {  // every origin's block of this generation must have arrived (as k_ex_wait)
        const unsigned long long t0 = vl_globaltimer();
        volatile unsigned long long *f = (volatile unsigned long long *)ex.in[tid].flag[par];
        while (*f < ex.seq) {
            if (vl_globaltimer() - t0 > timeout_ns) { raise(st, DE_EXCHANGE_TIMEOUT); break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
    __syncthreads();
    for (uint64_t c0 = (uint64_t)blockIdx.x * EXP_TPB; c0 < M; c0 += (uint64_t)gridDim.x * EXP_TPB) {
        const uint64_t idx = c0 + tid;
        const bool active = idx < M;
        uint32_t o = 0;
        uint64_t k = 0, slot = 0;
        uint4 meta = make_uint4(MIG_WILD_META, 0u, 0u, 0u);
        if (active) {
            o = ex_find(sc->imp_start, ex.world, idx);
            k = idx - sc->imp_start[o];
            slot = sc->in_start[o] + k;
            meta = __ldcg(ex.in[o].meta[par] + k);
        }
        const uint32_t is_rec = (active && meta.x != MIG_WILD_META) ? 1u : 0u;
        uint32_t words = is_rec ? meta.y : 0u;  // arena words: the base list (the delta entries go into the head)
        const uint32_t nd = is_rec ? (meta.z & 0xFFu) : 0u;
        if (is_rec && ((uint64_t)meta.x + words + nd > ex.in[o].w_cap || nd > HEAD_DELTA)) { raise(st, DE_REPLAY); words = 0u; meta.y = 0u; meta.z = 0u; }
        uint32_t ir = is_rec, iw = words;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, ir, d), z = __shfl_up_sync(0xFFFFFFFFu, iw, d);
            if (lane >= (uint32_t)d) { ir += y; iw += z; }
        }
        __syncthreads();
        if (lane == 31) { s_wr[wid] = ir; s_ww[wid] = iw; }
        __syncthreads();
        uint32_t rpre = ir - is_rec, wpre = iw - words, rtot = 0, wtot = 0;
#pragma unroll
        for (int w = 0; w < EXP_TPB / 32; w++) {
            if (w < (int)wid) { rpre += s_wr[w]; wpre += s_ww[w]; }
            rtot += s_wr[w];
            wtot += s_ww[w];
        }
        if (tid == 0) {  // records and arena words of the round: one reservation each
            unsigned long long rb = rtot ? atomicAdd(&st->n_hap, (unsigned long long)rtot) : 0ull;
            unsigned long long wb = wtot ? atomicAdd(&st->arena_top, (unsigned long long)wtot) : 0ull;
            if (rb + rtot > st->cap_hap) { raise(st, DE_HAP_CAPACITY); rb = ~0ull; }
            if (wb + wtot > st->cap_arena) { raise(st, DE_ARENA_CAPACITY); rb = ~0ull; }
            s_rbase = rb;
            s_wbase = wb;
        }
        __syncthreads();
        const bool room = s_rbase != ~0ull;
        const unsigned long long wbase = s_wbase;
        s_pref[tid] = wpre;
        if (tid == 0) s_pref[EXP_TPB] = wtot;
        s_src[tid] = meta.x;
        s_org[tid] = o;
        uint32_t id = 0;
        if (is_rec && room) {
            id = (uint32_t)(s_rbase + rpre);
            for (uint32_t p = 0; p < P; p++) set_fitness(st, p, id, __ldcg(ex.in[o].raw[par] + k * P + p));
            HapHead hd;
            hd.off = wbase + wpre;
            hd.len = meta.y;
            hd.nd_mlen = meta.z;
            hd.raw0 = P ? st->h_raw[0][id] : 1.0;
            hd.fit0 = P ? st->h_fit[0][id] : 1.0;
            const uint32_t *dsrc = ex.in[o].entries[par] + meta.x + meta.y;  // the delta entries follow the base list
            const uint32_t ndh = meta.z & 0xFFu;
#pragma unroll
            for (uint32_t q = 0; q < HEAD_DELTA; q++) hd.d[q] = q < ndh ? __ldcg(dsrc + q) : 0xFFFFFFFFu;
            st->head[id] = hd;
        }
        uint32_t h = 0, nm = 0;
        if (active && room && slot < cap) {
            const double fit = pi0 >= 0 ? st->h_fit[pi0][id] : 1.0;
            st->hap_id_next[slot] = id;
            st->pop_fit_next[slot] = fit;  // the mapped reproductive fitness travels with the infectant from here on (vl_fused.cuh)
            if (prefill) infect_draw(st, ic, slot, fit, h, nm);
        }
        __syncthreads();
        if (room) {  // the round's base lists into the arena words just reserved: consecutive stores over the whole CTA
            uint32_t *__restrict__ out = st->arena + wbase;
            for (uint32_t g0 = tid; g0 < wtot; g0 += 4 * EXP_TPB) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t g = g0 + u * EXP_TPB;
                    if (g < wtot) {
                        const uint32_t j = ex_owner(s_pref, g), r = g - s_pref[j];
                        v[u] = __ldcg(ex.in[s_org[j]].entries[par] + s_src[j] + r);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (g0 + u * EXP_TPB < wtot) out[g0 + u * EXP_TPB] = v[u];
            }
        }
        if (prefill) {  // (one item per thread and round: the list cannot overflow)
            CtaList cl{s_items, s_clcnt, EXP_TPB};
            cl_push(cl, nm != 0u, make_uint4((uint32_t)slot, nm, h, id));
            cl_flush(st, cl, nullptr, 0ull);
        }
    }
}

// ---- one compartment sharded over the ranks ----------------------------------------------------
// Hosts are split into `world` equal ranges, rank g keeps the infectants that infect its hosts.  The reference draws
// N' = (min(W * dilution, max_population)) as usize parents iid proportional to offspring over the whole population and
// gives every new infectant an iid uniform host (src/simulation.rs:365-387,491-527).  A new infectant's rank is therefore
// uniform, its parent's rank is proportional to that rank's offspring total W_o, independently:
//     n[target][origin] ~ Multinomial(N'; W_origin / (world * W))           over all world^2 cells,
// and given n, origin o draws n[t][o] parents iid proportional to its own offspring for target t — exactly the
// transmission block with device-decided amounts.  Every rank evaluates the same multinomial from the same inputs
// (the gathered totals) with the same Philox stream, so no rank has to be told its row.
__global__ void k_shard_publish_total(DevState *st, ExArgs ex) {
    __shared__ uint64_t s_red[TPB / 32];
    const uint64_t nt = tile_count(st);
    uint64_t local = 0;
    for (uint64_t t = threadIdx.x; t < nt; t += TPB) local += st->tile_sum[t];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t total = 0;
        for (int w = 0; w < TPB / 32; w++) total += s_red[w];
        s_red[0] = total;
    }
    __syncthreads();
    const uint32_t t = threadIdx.x;
    if (t < ex.world) {
        *(volatile uint64_t *)ex.tot_out[t][ex.parity] = s_red[0];
        __threadfence_system();
        *(volatile unsigned long long *)ex.tot_flag_out[t][ex.parity] = ex.seq;
    }
}
// wait for every rank's total, then split the draws: plan_amount[t] = n[t][rank]
__global__ void k_shard_split(DevState *st, ExArgs ex, unsigned long long timeout_ns) {
    __shared__ uint64_t s_W[EX_MAX_WORLD];
    const uint32_t o = threadIdx.x;
    if (o < ex.world) {
        const unsigned long long t0 = vl_globaltimer();
        volatile unsigned long long *f = (volatile unsigned long long *)(ex.tot_flag_in[ex.parity] + o);
        while (*f < ex.seq) {
            if (vl_globaltimer() - t0 > timeout_ns) { raise(st, DE_EXCHANGE_TIMEOUT); break; }
            __nanosleep(200);
        }
        __threadfence_system();
        s_W[o] = ((volatile uint64_t *)ex.tot_in[ex.parity])[o];
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const uint32_t G = ex.world;
    uint64_t W = 0;
    for (uint32_t g = 0; g < G; g++) W += s_W[g];
    double target = (double)W * st->dilution;
    const double cap = (double)st->max_population;
    if (!(target <= cap)) target = cap;
    uint64_t left = f64_as_usize(target);
    st->offspring_total = W;
    st->target_size = target;
    if (W == 0) { raise(st, DE_ALL_ZERO_WEIGHTS); left = 0; }
    // sequential conditional binomials over the G*G cells in (target, origin) order; cell weight W_o (the factor 1/G is common)
    double wleft = (double)W * (double)G;
    for (uint32_t t = 0; t < G; t++)
        for (uint32_t g = 0; g < G; g++) {
            const double w = (double)s_W[g];
            uint64_t c = 0;
            if (left > 0 && w > 0.0) {
                if (!(w < wleft)) c = left;  // the last cell with weight takes what is left
                else {
                    Philox rng(ex.shard_seed, 0xFFFFFFFEu, st->generation, PH_SHARD, (uint64_t)t * G + g);
                    c = binomial(rng, left, w / wleft);
                }
            }
            left -= c;
            wleft -= w;
            if (g == ex.rank) st->plan_amount[t] = c;
        }
}

}  // namespace vl
