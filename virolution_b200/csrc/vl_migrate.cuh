// vl_migrate.cuh — moving populations between handles (compartments / GPUs) and exporting them.
//
// The reference shares one Rc/Arc haplotype tree between all compartments of a process, so transmission
// (src/runner.rs:401-422) only moves references.  Here every handle owns its arena, so a detached population
// that crosses handles is turned into a self-contained image: every distinct haplotype once (flattened
// entries + memoised raw fitness per provider) plus one local index per infectant.  The same image is what
// travels over NVLink between GPUs.  Export kernels back Haplotype::get_mutations / get_base for
// final.<i>.csv and the parity tests (src/core/haplotype.rs:402-466, src/runner.rs:137-152).
#pragma once
#include "vl_kernels.cuh"

namespace vl {

constexpr uint64_t PACK_MAGIC = 0x564c5041434b3031ull;  // "VLPACK01"

struct PackHeader {
    uint64_t magic;
    uint64_t n_pop, n_unique, n_words, n_providers;
    uint64_t off_idx, off_len, off_off, off_raw, off_entries;  // byte offsets from the image start
    uint64_t bytes;
    uint64_t _pad;
};

__host__ __device__ inline uint64_t align16(uint64_t x) { return (x + 15) & ~15ull; }
__host__ inline PackHeader make_pack_header(uint64_t n_pop, uint64_t n_unique, uint64_t n_words, uint64_t n_providers) {
    PackHeader h;
    h.magic = PACK_MAGIC;
    h.n_pop = n_pop; h.n_unique = n_unique; h.n_words = n_words; h.n_providers = n_providers;
    uint64_t o = align16(sizeof(PackHeader));
    h.off_idx = o; o = align16(o + n_pop * 4);
    h.off_len = o; o = align16(o + n_unique * 4);
    h.off_off = o; o = align16(o + n_unique * 8);
    h.off_raw = o; o = align16(o + n_providers * n_unique * 8);
    h.off_entries = o; o = align16(o + n_words * 4);
    h.bytes = o;
    h._pad = 0;
    return h;
}

// The distinct haplotypes of a detached population are found with the compaction's bit map (k_gc_clear / k_gc_mark /
// k_gc_words + the two scans, vl_kernels.cuh; the wildtype — id 0, present on every handle — is never shipped); every one
// is written once, flattened.
__global__ void __launch_bounds__(TPB) k_pack_copy(DevState *st, const uint32_t *__restrict__ ids, uint8_t *image, PackHeader hd) {
    const uint64_t nw = (st->n_hap + 31) >> 5;
    uint32_t *o_idx = reinterpret_cast<uint32_t *>(image + hd.off_idx);
    uint32_t *o_len = reinterpret_cast<uint32_t *>(image + hd.off_len);
    uint64_t *o_off = reinterpret_cast<uint64_t *>(image + hd.off_off);
    double *o_raw = reinterpret_cast<double *>(image + hd.off_raw);
    uint32_t *o_ent = reinterpret_cast<uint32_t *>(image + hd.off_entries);
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, span = (uint64_t)gridDim.x * blockDim.x;
    if (tid == 0) *reinterpret_cast<PackHeader *>(image) = hd;
    for (uint64_t w = tid; w < nw; w += span) {
        uint32_t bits = st->gc_bits[w];
        if (!bits) continue;
        uint32_t u = st->gc_cbase[w];
        uint64_t noff = st->gc_lbase[w];
        while (bits) {
            const uint32_t b = __ffs(bits) - 1;
            bits &= bits - 1;
            const uint64_t h = (w << 5) + b;
            const HapHead hh = st->head[h];
            const uint32_t len = flatten_record(st, hh, o_ent + noff);
            o_len[u] = len;
            o_off[u] = noff;
            for (uint32_t p = 0; p < st->n_providers; p++) o_raw[(uint64_t)p * hd.n_unique + u] = hap_raw(st, p, h);
            u += 1;
            noff += len;
        }
    }
    for (uint64_t i = tid; i < hd.n_pop; i += span) {
        const uint32_t id = ids[i];
        o_idx[i] = id == 0 ? NONE32 : gc_new_id(st, id);
    }
}

// reserve[0] = first new haplotype id (NONE32 on failure), reserve[1] = first arena word
__global__ void k_reserve_bulk(DevState *st, uint64_t n_records, uint64_t n_words, uint64_t *reserve) {
    const uint64_t id0 = st->n_hap, a0 = st->arena_top;
    if (id0 + n_records > st->cap_hap) { raise(st, DE_HAP_CAPACITY); reserve[0] = NONE32; return; }
    if (a0 + n_words > st->cap_arena) { raise(st, DE_ARENA_CAPACITY); reserve[0] = NONE32; return; }
    st->n_hap = id0 + n_records;
    st->arena_top = a0 + n_words;
    reserve[0] = id0;
    reserve[1] = a0;
}
// dst == nullptr: write into the population being assembled at dst_off
__global__ void __launch_bounds__(TPB) k_unpack(DevState *st, const uint8_t *__restrict__ image, PackHeader hd, const uint64_t *reserve,
                                                uint32_t *dst, uint64_t dst_off) {
    const uint32_t *i_idx = reinterpret_cast<const uint32_t *>(image + hd.off_idx);
    const uint32_t *i_len = reinterpret_cast<const uint32_t *>(image + hd.off_len);
    const uint64_t *i_off = reinterpret_cast<const uint64_t *>(image + hd.off_off);
    const double *i_raw = reinterpret_cast<const double *>(image + hd.off_raw);
    const uint32_t *i_ent = reinterpret_cast<const uint32_t *>(image + hd.off_entries);
    if (dst == nullptr) dst = st->hap_id_next;
    dst += dst_off;
    const uint64_t id0 = reserve[0], a0 = reserve[1];
    const bool ok = id0 != NONE32;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, span = (uint64_t)gridDim.x * blockDim.x;
    if (ok) {
        for (uint64_t u = tid; u < hd.n_unique; u += span) {
            const uint64_t id = id0 + u;
            for (uint32_t p = 0; p < st->n_providers; p++) set_fitness(st, p, id, i_raw[(uint64_t)p * hd.n_unique + u]);
            write_flat_head(st, id, a0 + i_off[u], i_len[u]);
        }
        for (uint64_t w = tid; w < hd.n_words; w += span) st->arena[a0 + w] = i_ent[w];
    }
    for (uint64_t i = tid; i < hd.n_pop; i += span) {
        const uint32_t u = i_idx[i];
        dst[i] = (u == NONE32 || !ok) ? 0u : (uint32_t)(id0 + u);
    }
}

// ---- migrants between processes (one process per GPU): per-migrant images, no dedup --------------------------
// The image of a set of detached populations is three flat arrays over the concatenated migrants:
//   len[m]      entries of migrant m's haplotype, MIG_WILDTYPE for the wildtype (exists on every handle)
//   raw[m*P+p]  memoised raw fitness under provider p (AttributeSet, src/core/attributes.rs:196-217)
//   entries[]   the flattened records, migrant after migrant
// so the part destined for one target is a contiguous range of each array — what an all-to-all-v wants.  The
// reference shares one haplotype tree between compartments and moves references (src/runner.rs:401-422); every
// migrant gets a record of its own on the target, like every mutant gets its own node there.
constexpr uint32_t MIG_WILDTYPE = 0x80000000u;
__global__ void __launch_bounds__(TPB) k_mig_lens(DevState *st, const uint32_t *__restrict__ ids, uint64_t m, uint64_t m0,
                                                  uint32_t *__restrict__ len_out, uint32_t *__restrict__ words_out, double *__restrict__ raw_out) {
    const uint32_t P = st->n_providers;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t id = ids[i];
        const uint32_t len = id == 0 ? 0u : st->head[id].nd_mlen >> 8;  // entries of the flattened record
        len_out[m0 + i] = id == 0 ? MIG_WILDTYPE : len;
        words_out[m0 + i] = len;
        for (uint32_t p = 0; p < P; p++) raw_out[(m0 + i) * P + p] = hap_raw(st, p, id);
    }
}
__global__ void __launch_bounds__(TPB) k_mig_copy(DevState *st, const uint32_t *__restrict__ ids, uint64_t m, uint64_t m0,
                                                  const uint64_t *__restrict__ word_off, uint32_t *__restrict__ entries_out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t id = ids[i];
        if (id == 0) continue;
        const HapHead hh = st->head[id];
        flatten_record(st, hh, entries_out + word_off[m0 + i]);
    }
}
// words per migrant of a received image (the wildtype marker carries no words)
__global__ void __launch_bounds__(TPB) k_mig_words(const uint32_t *__restrict__ len_in, uint64_t m, uint32_t *__restrict__ words_out,
                                                   uint32_t *__restrict__ rec_out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t l = len_in[i];
        words_out[i] = l == MIG_WILDTYPE ? 0u : l;
        rec_out[i] = l == MIG_WILDTYPE ? 0u : 1u;
    }
}
// totals[0] = records needed, totals[1] = words needed (from the two scans); reserve[0] = first id (NONE32 on failure)
__global__ void k_mig_reserve(DevState *st, const uint64_t *totals, uint64_t n_words_expected, uint64_t *reserve) {
    const uint64_t id0 = st->n_hap, a0 = st->arena_top;
    reserve[0] = NONE32;
    if (totals[1] != n_words_expected) { raise(st, DE_REPLAY); return; }
    if (id0 + totals[0] > st->cap_hap) { raise(st, DE_HAP_CAPACITY); return; }
    if (a0 + totals[1] > st->cap_arena) { raise(st, DE_ARENA_CAPACITY); return; }
    st->n_hap = id0 + totals[0];
    st->arena_top = a0 + totals[1];
    reserve[0] = id0;
    reserve[1] = a0;
}
__global__ void __launch_bounds__(TPB) k_mig_import(DevState *st, const uint32_t *__restrict__ len_in, const double *__restrict__ raw_in,
                                                    const uint32_t *__restrict__ entries_in, uint64_t m, const uint32_t *__restrict__ rec_off,
                                                    const uint64_t *__restrict__ word_off, const uint64_t *reserve, uint32_t *__restrict__ ids_out) {
    const uint64_t id0 = reserve[0], a0 = reserve[1];
    const bool ok = id0 != NONE32;
    const uint32_t P = st->n_providers;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t l = len_in[i];
        if (l == MIG_WILDTYPE || !ok) { ids_out[i] = 0u; continue; }
        const uint64_t id = id0 + rec_off[i], off = a0 + word_off[i];
        for (uint32_t p = 0; p < P; p++) set_fitness(st, p, id, raw_in[i * P + p]);
        const uint32_t *src = entries_in + word_off[i];
        uint32_t *dst = st->arena + off;
        for (uint32_t k = 0; k < l; k++) dst[k] = src[k];
        write_flat_head(st, id, off, l);
        ids_out[i] = (uint32_t)id;
    }
}

// population assembly from parts that already live on this handle
__global__ void __launch_bounds__(TPB) k_copy_ids(DevState *st, const uint32_t *__restrict__ src, uint64_t m, uint64_t dst_off) {
    uint32_t *dst = st->hap_id_next + dst_off;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_set_next_n(DevState *st, uint64_t n) { st->n_next = n; }
__global__ void k_fill_population(DevState *st, uint64_t n, uint32_t id) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) st->hap_id[i] = id;
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->n = n; st->n_sorted = 0; }
}
// copy of the first n_next entries of the assembled population (detached subsample of the fused path)
__global__ void __launch_bounds__(TPB) k_copy_out_next(DevState *st, uint32_t *__restrict__ dst, uint64_t cap) {
    const uint64_t m = st->n_next < cap ? st->n_next : cap;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = st->hap_id_next[i];
}

// target_size over the tile totals (src/simulation.rs:491-503); does not raise on an all-zero map
__global__ void __launch_bounds__(TPB) k_target_size(DevState *st) {
    __shared__ uint64_t s_red[TPB / 32];
    const uint64_t n = st->n;
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
        double target = 0.0;
        if (n > 0) {
            double a = (double)total * st->dilution, b = (double)st->max_population;
            target = a <= b ? a : b;
        }
        st->offspring_total = total;
        st->target_size = target;
    }
}

// ---- export: Haplotype::get_mutations per infectant (entries whose mutation-set view is present)
__global__ void __launch_bounds__(TPB) k_export_count(DevState *st, uint32_t *__restrict__ cnt) {
    const uint64_t n = st->n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const HapHead hh = st->head[st->hap_id[i]];
        uint32_t c = 0;
        merged_for_each(st->arena + hh.off, hh.len, hh.d, head_nd(hh), [&](uint32_t e) { c += entry_m(e) < 4 ? 1u : 0u; });
        cnt[i] = c;
    }
}
__global__ void __launch_bounds__(TPB) k_export_fill(DevState *st, const uint64_t *__restrict__ offsets, uint32_t *__restrict__ pos_out,
                                                     uint8_t *__restrict__ base_out) {
    const uint64_t n = st->n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const HapHead hh = st->head[st->hap_id[i]];
        uint64_t o = offsets[i];
        merged_for_each(st->arena + hh.off, hh.len, hh.d, head_nd(hh), [&](uint32_t e) {
            if (entry_m(e) < 4) { pos_out[o] = entry_pos(e); base_out[o] = (uint8_t)entry_m(e); o++; }
        });
    }
}
// ---- population statistics (src/stats/population.rs) ---------------------------------------------------------------
// PopulationFrequencies::frequencies (:15-40): per (site, base) the number of infectants whose mutation set holds that
// base there; the wildtype column is filled in by the caller as n minus the row sum, like the reference does
__global__ void __launch_bounds__(TPB) k_stats_counts(DevState *st, uint32_t *__restrict__ counts) {
    const uint64_t n = st->n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const HapHead hh = st->head[st->hap_id[i]];
        merged_for_each(st->arena + hh.off, hh.len, hh.d, head_nd(hh), [&](uint32_t e) {
            const uint32_t m = entry_m(e);
            if (m < 4) atomicAdd(&counts[(uint64_t)entry_pos(e) * 4 + m], 1u);
        });
    }
}
// PopulationStatistics::mean (:50-56): sum of the mapped attribute over infectants; one partial sum per CTA in a fixed
// tree order (the caller adds the partials in order), so the result does not depend on scheduling
__global__ void __launch_bounds__(TPB) k_stats_fitness_sum(DevState *st, int provider, double *__restrict__ partial) {
    __shared__ double s_red[TPB];
    const uint64_t n = st->n;
    const uint64_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per * blockIdx.x < n ? per * blockIdx.x : n, hi = lo + per < n ? lo + per : n;
    double acc = 0.0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += TPB) acc += provider >= 0 ? st->h_fit[provider][st->hap_id[i]] : 1.0;
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = TPB / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s_red[0];
}
__global__ void k_get_base(DevState *st, uint64_t infectant, uint32_t pos, uint32_t *out) {
    if (infectant >= st->n || pos >= st->L) { *out = 0xFFu; return; }
    const HapHead hh = st->head[st->hap_id[infectant]];
    const uint32_t e = merged_find(st->arena + hh.off, hh.len, hh.d, head_nd(hh), pos, st->L);
    *out = e != 0xFFFFFFFFu ? entry_g(e) : st->wildtype[pos];  // Haplotype::get_base (src/core/haplotype.rs:402-408)
}

}  // namespace vl
