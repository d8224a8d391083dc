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

namespace vl {

constexpr uint32_t EX_MAX_WORLD = MAX_PARTS;
constexpr uint64_t EX_HEADER_BYTES = 64;

// where one (target <- origin) block lives, for both parities; pointers are valid on the GPU that dereferences them
struct ExBlock {
    uint64_t *header[2];  // {m, w}
    uint32_t *len[2];
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
