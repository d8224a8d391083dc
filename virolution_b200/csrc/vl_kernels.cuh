// vl_kernels.cuh — the per-generation population update as sm_100a kernels.
//
// Kernel                reference rows it replaces
//   k_infect            BasicSimulation::infect + BasicHost::infect          src/simulation.rs:365-387,56-65
//   k_scatter_hosts,    HostMapBuffer::build (counting sort, back-fill)       src/core/hosts.rs:166-192
//   k_sort_slices[_heavy]   ... including the descending in-host order its back-fill produces
//   k_recombine         BasicHost::mutate, recombination part                 src/simulation.rs:68-90
//   k_mutation_counts,  BasicHost::mutate, mutation part + create_descendant  src/simulation.rs:93-117,
//   k_mutate                + FitnessProvider::get_fitness for every provider  src/core/haplotype.rs:241-260
//   k_fitness,k_host_sum[_heavy],k_offspring   BasicHost::replicate           src/simulation.rs:120-154,430-485
//   k_plan_resample     target_size + sample size                             src/simulation.rs:491-527
//   k_resample          Population::sample / sample_indices / select          src/core/population.rs:385-430
//   k_gc_*              Rc/Arc drop of unreachable haplotypes                 src/references/sync.rs:107-115
// All kernels are grid-stride over device-resident sizes (DevState::n etc.), so a whole generation is
// enqueued without the host ever reading a size back.
#pragma once
#include "vl_rng.cuh"
#include "vl_scan.cuh"
#include "vl_state.cuh"

namespace vl {

constexpr int TPB = 256;
constexpr uint32_t THREAD_TIER = 32;      // slices up to this length are handled by one thread
constexpr uint32_t SMEM_SORT_MAX = 8192;  // slices up to this length are sorted in shared memory

__device__ __forceinline__ void raise(DevState *st, uint32_t bit) { atomicOr(&st->error, bit); }

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
}

// ------------------------------------------------------------------------------------------ infect
// replay_inf: recorded Option<usize> per infectant (-1 = None) or nullptr.
__global__ void __launch_bounds__(TPB) k_infect(DevState *st, uint32_t *__restrict__ count, const int64_t *__restrict__ replay_inf,
                                                int build_csr) {
    const uint64_t n = st->n;
    const uint64_t H = st->H;
    const uint32_t gen = st->generation;
    const uint32_t *__restrict__ hap_id = st->hap_id;
    uint32_t *__restrict__ host_id = st->host_id;
    uint32_t *__restrict__ rank = st->rank;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->infectant_generations += n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t h = NONE32;
        if (replay_inf) {
            int64_t r = replay_inf[i];
            if (r >= 0) {
                if ((uint64_t)r < H) h = (uint32_t)r; else raise(st, DE_REPLAY);
            }
        } else if (H > 0) {
            Philox g(st->seed, st->compartment, gen, PH_INFECT, i);
            for (uint64_t hit = 0; hit < st->n_hits; hit++) {
                uint64_t cand = g.below(H);
                int s = find_spec(st, cand);
                if (s < 0) continue;
                double infectivity = 1.;
                int pi = st->specs[s].infect;
                if (pi >= 0) infectivity = utility_apply(st->prov[pi], st->h_raw[pi][hap_id[i]]);
                if (g.u01() < infectivity) { h = (uint32_t)cand; break; }
            }
        }
        host_id[i] = h;
        if (build_csr && h != NONE32) rank[i] = atomicAdd(&count[h], 1u);
    }
}

__global__ void __launch_bounds__(TPB) k_scatter_hosts(DevState *st) {
    const uint64_t n = st->n;
    const uint32_t *__restrict__ host_id = st->host_id;
    const uint32_t *__restrict__ rank = st->rank;
    const uint32_t *__restrict__ offsets = st->offsets;
    uint32_t *__restrict__ hosts = st->hosts;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t h = host_id[i];
        if (h != NONE32) hosts[offsets[h] + rank[i]] = (uint32_t)i;
    }
}

// The reference back-fills each host's slice while walking infectants upwards, which leaves the slice in
// DESCENDING infectant order (KAT src/core/hosts.rs:317-318).  Slices arrive here in arrival order of the
// atomics; sort them.  Short slices: one thread, insertion sort.  Long ones go to the heavy worklist.
__global__ void __launch_bounds__(TPB) k_sort_slices(DevState *st) {
    const uint64_t H = st->H;
    const uint32_t *__restrict__ offsets = st->offsets;
    uint32_t *__restrict__ hosts = st->hosts;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < H; h += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t b = offsets[h], e = offsets[h + 1];
        uint32_t len = e - b;
        if (len <= 1) continue;
        if (len > THREAD_TIER) {
            uint32_t slot = atomicAdd(&st->n_heavy, 1u);
            st->heavy_hosts[slot] = (uint32_t)h;
            continue;
        }
        for (uint32_t a = b + 1; a < e; a++) {
            uint32_t x = hosts[a];
            uint32_t c = a;
            while (c > b && hosts[c - 1] < x) { hosts[c] = hosts[c - 1]; c--; }
            hosts[c] = x;
        }
    }
}

// one CTA per heavy slice: normalised bitonic network (all comparators put the larger index first, missing
// partners beyond the slice end are no-ops), in shared memory when the slice fits, else in global memory.
__global__ void __launch_bounds__(TPB) k_sort_slices_heavy(DevState *st) {
    extern __shared__ uint32_t s_keys[];
    const uint32_t n_heavy = st->n_heavy;
    for (uint32_t w = blockIdx.x; w < n_heavy; w += gridDim.x) {
        const uint32_t h = st->heavy_hosts[w];
        const uint32_t b = st->offsets[h], len = st->offsets[h + 1] - b;
        uint32_t *g = st->hosts + b;
        const bool in_smem = len <= SMEM_SORT_MAX;
        uint32_t *a = in_smem ? s_keys : g;
        if (in_smem) {
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) s_keys[i] = g[i];
        }
        __syncthreads();
        uint32_t P = 1;
        while (P < len) P <<= 1;
        for (uint32_t k = 2; k <= P; k <<= 1) {
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
                uint32_t l = i ^ (k - 1);
                if (l > i && l < len) {
                    uint32_t x = a[i], y = a[l];
                    if (x < y) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
            for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
                    uint32_t l = i ^ j;
                    if (l > i && l < len) {
                        uint32_t x = a[i], y = a[l];
                        if (x < y) { a[i] = y; a[l] = x; }
                    }
                }
                __syncthreads();
            }
        }
        if (in_smem) {
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) g[i] = s_keys[i];
        }
        __syncthreads();
    }
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
__device__ __forceinline__ uint32_t ch_from(uint32_t c) { return (c >> 4) & 0xF; }
__device__ __forceinline__ uint32_t ch_to(uint32_t c) { return c & 0xF; }

// EpistasisTable::update_fitness (src/core/fitness/epistasis.rs:105-173) on a freshly merged record.
// Quirks kept: candidates are the changes whose (pos,to) has a row (`from` is never tested); entries of the
// mutation set at positions <= the change that belong to this node's own changes are skipped.
__device__ inline double epi_update(const DevProvider &p, const uint32_t *chs, uint32_t n_ch, const uint32_t *ent, uint32_t n_ent) {
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
        for (uint32_t m = 0; m < n_ent; m++) {
            const uint32_t mb = entry_m(ent[m]);
            if (mb >= 4) continue;
            const uint32_t pos = entry_pos(ent[m]);
            if (pos <= cpos) {
                bool own = false;
                for (uint32_t q = 0; q < n_ch; q++) if (ch_pos(chs[q]) == pos) { own = true; break; }
                if (own) continue;
            }
            if (pos != cpos) {
                double v;
                if (epi_get(p, alo, ahi, ekey(pos, mb), v)) fitness *= v;
                if (rhi > rlo && epi_get(p, rlo, rhi, ekey(pos, mb), v)) fitness /= v;
            }
        }
    }
    return fitness;
}

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
    const uint32_t ll = st->h_len[left], lr = st->h_len[right];
    NewRecord r = reserve_record_single(st, ll + lr);
    if (!r.ok) return left;
    const uint32_t *el = st->arena + st->h_off[left], *er = st->arena + st->h_off[right];
    uint32_t *out = st->arena + r.off;
    uint32_t o = 0;
    uint32_t k = 0;
    for (; k < lr && entry_pos(er[k]) < s0; k++) out[o++] = er[k];
    for (uint32_t q = 0; q < ll; q++) {
        uint32_t p = entry_pos(el[q]);
        if (p >= s0 && p < s1) out[o++] = el[q];
    }
    for (; k < lr; k++) if (entry_pos(er[k]) >= s1) out[o++] = er[k];
    st->h_off[r.id] = r.off;
    st->h_len[r.id] = o;
    st->h_parent[r.id] = left;
    for (uint32_t p = 0; p < st->n_providers; p++) st->h_raw[p][r.id] = compute_fitness_full(st->prov[p], out, o);
    return r.id;
}

__global__ void __launch_bounds__(TPB) k_recombine(DevState *st, ReplayRecomb rp, int replay) {
    const uint64_t H = st->H;
    const uint32_t L = st->L;
    const double rho = st->host_recombination_rate;
    const uint32_t *__restrict__ offsets = st->offsets;
    const uint32_t *__restrict__ hosts = st->hosts;
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
                uint32_t ii = hosts[b + i], jj = hosts[b + j];
                uint32_t rec = make_recombinant(st, hap_id[ii], hap_id[jj], rp.s0[e], rp.s1[e]);
                hap_id[rp.which[e] ? jj : ii] = rec;
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
            uint32_t ii = hosts[b + i], jj = hosts[b + j];
            uint32_t rec = make_recombinant(st, hap_id[ii], hap_id[jj], s0, s1);
            uint32_t which = (uint32_t)g.below(2);
            hap_id[which ? jj : ii] = rec;
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

// n_mut ~ Binomial(L, mu) per infected infectant (src/simulation.rs:94-95); the ones with n_mut > 0 go
// to a worklist so that the expensive part runs on dense warps.
__global__ void __launch_bounds__(TPB) k_mutation_counts(DevState *st, ReplayMut rp, int replay) {
    __shared__ uint32_t s_base;
    const uint64_t n = st->n;
    const uint32_t *__restrict__ host_id = st->host_id;
    const uint64_t L = st->L;
    const double mu = st->host_mutation_rate;
    const uint64_t span = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + span - 1) / span;
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t i = r * span + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t nm = 0;
        if (i < n && host_id[i] != NONE32) {
            if (replay) {
                nm = (uint32_t)(rp.offsets[i + 1] - rp.offsets[i]);
            } else {
                Philox g(st->seed, st->compartment, st->generation, PH_MUTCOUNT, i);
                nm = (uint32_t)binomial(g, L, mu);
            }
        } else if (i < n && replay && rp.offsets[i + 1] != rp.offsets[i]) {
            raise(st, DE_REPLAY);  // the reference never mutates an uninfected infectant
        }
        uint32_t tot;
        uint32_t off = block_excl_scan(nm ? 1u : 0u, &tot);
        if (threadIdx.x == 0 && tot) s_base = atomicAdd(&st->n_mut_list, tot);
        __syncthreads();
        if (nm) {
            st->mut_list[s_base + off] = (uint32_t)i;
            st->mut_count[s_base + off] = nm;
        }
        __syncthreads();
    }
}

// One thread per mutated infectant: draw sites (with replacement) and sort them, look up the current base
// on the pre-mutation haplotype, draw the substitution, merge into a new flattened record and update the
// raw fitness of every provider incrementally: raw(parent) * fold(acc * T[pos,to] / T[pos,from])
// (src/core/fitness/table.rs:65-68) [* epistatic update].
__global__ void __launch_bounds__(TPB) k_mutate(DevState *st, ReplayMut rp, int replay) {
    const uint32_t n_list = st->n_mut_list;
    const uint32_t L = st->L;
    const uint8_t *__restrict__ wt = st->wildtype;
    const uint32_t span = gridDim.x * blockDim.x;
    const uint32_t rounds = (n_list + span - 1) / span;
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t w = r * span + blockIdx.x * blockDim.x + threadIdx.x;
        const bool active = w < n_list;
        uint32_t i = 0, nm = 0, pid = 0, plen = 0;
        if (active) {
            i = st->mut_list[w];
            nm = st->mut_count[w];
            pid = st->hap_id[i];
            plen = st->h_len[pid];
        }
        NewRecord rec = reserve_record(st, active, plen + 2 * nm);
        if (!rec.ok) continue;
        const uint32_t *pe = st->arena + st->h_off[pid];
        uint32_t *out = st->arena + rec.off;
        uint32_t *chs = out + plen + nm;
        if (replay) {
            const uint64_t ro = rp.offsets[i];
            for (uint32_t k = 0; k < nm; k++) {
                uint32_t site = rp.sites[ro + k];
                if (site >= L || (k > 0 && site < ch_pos(chs[k - 1])) || rp.bases[ro + k] > 3) { raise(st, DE_REPLAY); site = site % L; }
                uint32_t cur = entries_get_base(pe, plen, site, wt);
                uint32_t to = rp.bases[ro + k] & 3;
                if (cur == to) raise(st, DE_SAME_BASE);
                chs[k] = (site << 8) | (cur << 4) | to;
            }
        } else {
            Philox g(st->seed, st->compartment, st->generation, PH_MUTATE, i);
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
                uint32_t site = chs[k];
                uint32_t cur = entries_get_base(pe, plen, site, wt);
                uint32_t to = (uint32_t)weighted4(g, &st->host_subst[cur * 4]);
                if (cur == to) raise(st, DE_SAME_BASE);
                chs[k] = (site << 8) | (cur << 4) | to;
            }
        }
        // merge parent entries with the change groups
        uint32_t a = 0, b = 0, o = 0;
        while (a < plen || b < nm) {
            uint32_t pa = a < plen ? entry_pos(pe[a]) : 0xFFFFFFFFu;
            uint32_t pb = b < nm ? ch_pos(chs[b]) : 0xFFFFFFFFu;
            if (pa < pb) {
                out[o++] = pe[a++];
            } else {
                const uint32_t first_to = ch_to(chs[b]);
                uint32_t last_to = first_to;
                while (b < nm && ch_pos(chs[b]) == pb) { last_to = ch_to(chs[b]); b++; }
                if (pa == pb) a++;
                const uint32_t w0 = wt[pb];
                const uint32_t m = (last_to == w0) ? 4u : last_to;
                if (!(first_to == w0 && m == 4u)) out[o++] = make_entry(pb, first_to, m);
            }
        }
        st->h_off[rec.id] = rec.off;
        st->h_len[rec.id] = o;
        st->h_parent[rec.id] = pid;
        for (uint32_t p = 0; p < st->n_providers; p++) {
            const DevProvider &pr = st->prov[p];
            double raw = 1.0;
            if (pr.kind != VL_FITNESS_NEUTRAL) {
                double acc = 1.;
                for (uint32_t k = 0; k < nm; k++) {
                    const uint64_t row = (uint64_t)ch_pos(chs[k]) * 4;
                    acc = acc * pr.table[row + ch_to(chs[k])] / pr.table[row + ch_from(chs[k])];
                }
                if (pr.kind == VL_FITNESS_EPISTATIC) acc = acc * epi_update(pr, chs, nm, out, o);
                raw = st->h_raw[p][pid] * acc;
            }
            st->h_raw[p][rec.id] = raw;
        }
        st->hap_id[i] = rec.id;
    }
}

// ------------------------------------------------------------------------------------------ replication
// mapped fitness of every infectant under its host's reproductive provider (or 1.0), NaN where uninfected
__global__ void __launch_bounds__(TPB) k_fitness(DevState *st, int atomic_sum) {
    const uint64_t n = st->n;
    const uint32_t *__restrict__ host_id = st->host_id;
    const uint32_t *__restrict__ hap_id = st->hap_id;
    double *__restrict__ fit = st->fit;
    const bool single = st->n_specs == 1 && st->specs[0].lo == 0 && st->specs[0].hi >= st->H;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t h = host_id[i];
        double f = __longlong_as_double(0x7FF8000000000000ll);
        if (h != NONE32) {
            const int s = single ? 0 : find_spec(st, h);
            const int pi = s >= 0 ? st->specs[s].repro : -1;
            f = pi >= 0 ? utility_apply(st->prov[pi], st->h_raw[pi][hap_id[i]]) : 1.;
            if (atomic_sum) atomicAdd(&st->hsum[h], f);
        }
        fit[i] = f;
    }
}

// S_h = sum of f over the host's slice in slice order, from 0.0 (iter().sum(), src/simulation.rs:144)
__global__ void __launch_bounds__(TPB) k_host_sum(DevState *st) {
    const uint64_t H = st->H;
    const uint32_t *__restrict__ offsets = st->offsets;
    const uint32_t *__restrict__ hosts = st->hosts;
    const double *__restrict__ fit = st->fit;
    double *__restrict__ hsum = st->hsum;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < H; h += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = offsets[h], e = offsets[h + 1];
        if (e - b > THREAD_TIER) continue;  // heavy tier
        if (e == b) continue;
        double s = 0.0;
        for (uint32_t a = b; a < e; a++) s += fit[hosts[a]];
        hsum[h] = s;
    }
}
// one warp per heavy host: lanes gather 32 terms at a time, the additions stay strictly sequential
__global__ void __launch_bounds__(TPB) k_host_sum_heavy(DevState *st) {
    const uint32_t n_heavy = st->n_heavy;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = warp; w < n_heavy; w += n_warps) {
        const uint32_t h = st->heavy_hosts[w];
        const uint32_t b = st->offsets[h], e = st->offsets[h + 1];
        double s = 0.0;
        for (uint32_t a0 = b; a0 < e; a0 += 32) {
            double f = (a0 + lane < e) ? st->fit[st->hosts[a0 + lane]] : 0.0;
            const uint32_t cnt = min(32u, e - a0);
            for (uint32_t t = 0; t < cnt; t++) s += __shfl_sync(0xFFFFFFFFu, f, t);
        }
        if (lane == 0) st->hsum[h] = s;
    }
}

// offspring_i ~ Poisson(f_i * R0 / S_host) (src/simulation.rs:146-152); one CTA per resample tile so the
// tile's offspring total comes out of the same pass.
__global__ void __launch_bounds__(TPB) k_offspring(DevState *st, const uint64_t *__restrict__ replay_off) {
    __shared__ uint64_t s_red[TPB / 32];
    const uint64_t n = st->n;
    const uint64_t n_tiles = (n + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
    const uint32_t *__restrict__ host_id = st->host_id;
    const double *__restrict__ fit = st->fit;
    const double *__restrict__ hsum = st->hsum;
    uint32_t *__restrict__ offspring = st->offspring;
    const double R0 = st->host_r0;
    const uint32_t gen = st->generation;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t local = 0;
#pragma unroll 1
        for (uint32_t k = 0; k < RESAMPLE_TILE / TPB; k++) {
            const uint64_t i = tile * RESAMPLE_TILE + (uint64_t)k * TPB + threadIdx.x;
            if (i >= n) break;
            const uint32_t h = host_id[i];
            uint32_t off = 0;
            if (replay_off) {
                uint64_t v = replay_off[i];
                if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0; }
                off = (uint32_t)v;
            } else if (h != NONE32) {
                const double f = fit[i];
                const double lam = f * R0 / hsum[h];
                if (isfinite(lam) && lam > 0.0 && lam <= 1.844e19) {
                    Philox g(st->seed, st->compartment, gen, PH_OFFSPRING, i);
                    uint64_t v = poisson(g, lam);
                    if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0xFFFFFFFFull; }
                    off = (uint32_t)v;
                } else if (__double_as_longlong(f) != 0ll) {
                    // the reference leaves the f64 bit pattern of f in the slot (src/simulation.rs:146-152):
                    // undefined population dynamics — report instead of imitating
                    raise(st, DE_UNDEFINED_OFFSPRING);
                }
            }
            offspring[i] = off;
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
// tile sums of an offspring map that came from the host (vl_target_size / vl_sample_indices with an explicit map)
__global__ void __launch_bounds__(TPB) k_tile_sums(DevState *st) {
    __shared__ uint64_t s_red[TPB / 32];
    const uint64_t n = st->n;
    const uint64_t n_tiles = (n + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t local = 0;
        for (uint32_t k = 0; k < RESAMPLE_TILE / TPB; k++) {
            const uint64_t i = tile * RESAMPLE_TILE + (uint64_t)k * TPB + threadIdx.x;
            if (i < n) local += st->offspring[i];
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
// One CTA.  (1) segment tree over the tile totals; (2) target_size = min(sum*dilution, max_population)
// (src/simulation.rs:491-503) and size = (factor*target) as usize (:521) or an explicit amount
// (sample_indices, :505-514); (3) the `size` iid draws proportional to offspring (Population::sample,
// src/core/population.rs:385-403) are split over tiles as an exact multinomial by walking the tree with
// conditional binomials; (4) exclusive scan of the per-tile draw counts.
__global__ void __launch_bounds__(PLAN_THREADS) k_plan_resample(DevState *st, double factor, int use_amount, uint64_t amount,
                                                               int set_next, uint32_t salt) {
    __shared__ uint64_t s_scan[PLAN_THREADS / 32];
    __shared__ uint64_t s_carry;
    const uint64_t n = st->n;
    const uint64_t nt = (n + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
    const uint64_t P = pow2_ceil(nt ? nt : 1);
    uint64_t *tree = st->tile_tree, *cnt = st->tile_cnt;
    const uint32_t tid = threadIdx.x;
    for (uint64_t t = tid; t < P; t += PLAN_THREADS) tree[P + t] = t < nt ? st->tile_sum[t] : 0;
    __syncthreads();
    for (uint64_t s = P >> 1; s >= 1; s >>= 1) {
        for (uint64_t node = s + tid; node < 2 * s; node += PLAN_THREADS) tree[node] = tree[2 * node] + tree[2 * node + 1];
        __syncthreads();
    }
    if (tid == 0) {
        const uint64_t total = tree[1];
        double target = 0.0;
        if (n > 0) {
            double a = (double)total * st->dilution, b = (double)st->max_population;
            target = a <= b ? a : b;
        }
        uint64_t size = use_amount ? amount : f64_as_usize(factor * target);
        if (n == 0) size = 0;
        if (n > 0 && total == 0) { raise(st, DE_ALL_ZERO_WEIGHTS); size = 0; }
        if (set_next && size > st->cap_inf) { raise(st, DE_INF_CAPACITY); size = st->cap_inf; }
        st->offspring_total = total;
        st->target_size = target;
        st->n_next = size;
        cnt[1] = size;
    }
    __syncthreads();
    for (uint64_t s = 1; s < P; s <<= 1) {
        for (uint64_t node = s + tid; node < 2 * s; node += PLAN_THREADS) {
            const uint64_t c = cnt[node], wl = tree[2 * node], w = tree[node];
            uint64_t left = 0;
            if (c > 0 && wl > 0) {
                if (wl == w) left = c;
                else {
                    Philox g(st->seed, st->compartment, st->generation, PH_PLAN, node | ((uint64_t)salt << 48));
                    left = binomial(g, c, (double)wl / (double)w);
                }
            }
            cnt[2 * node] = left;
            cnt[2 * node + 1] = c - left;
        }
        __syncthreads();
    }
    // exclusive scan of the leaf counts
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t t0 = 0; t0 < nt; t0 += PLAN_THREADS) {
        const uint64_t t = t0 + tid;
        const uint64_t x = t < nt ? cnt[P + t] : 0;
        uint64_t inc = x;
        const uint32_t lane = tid & 31, wid = tid >> 5;
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
        const uint64_t carry = s_carry;
        if (t < nt) st->tile_base[t] = carry + woff + inc - x;
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
        __syncthreads();
    }
}

// One CTA per tile: stage the tile's offspring (as an inclusive CDF) and haplotype ids in shared memory,
// then every draw is a uniform in [0, tile total) + a binary search in shared memory + one coalesced
// store.  No global gathers.  mode 0: write haplotype ids of the drawn infectants to dst; mode 1: write
// the drawn infectant indices (u64) to dst64.
__global__ void __launch_bounds__(TPB) k_resample(DevState *st, uint32_t *__restrict__ dst, uint64_t *__restrict__ dst64, int mode,
                                                  uint32_t stream_salt) {
    __shared__ uint64_t s_cdf[RESAMPLE_TILE];
    __shared__ uint32_t s_hap[RESAMPLE_TILE];
    __shared__ uint64_t s_w[TPB / 32];
    const uint64_t n = st->n;
    const uint64_t nt = (n + RESAMPLE_TILE - 1) / RESAMPLE_TILE;
    const uint64_t P = pow2_ceil(nt ? nt : 1);
    const uint32_t *__restrict__ offspring = st->offspring;
    const uint32_t *__restrict__ hap_id = st->hap_id;
    if (dst == nullptr && mode == 0) dst = st->hap_id_next;
    constexpr uint32_t PER = RESAMPLE_TILE / TPB;  // consecutive items per thread
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (uint64_t tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        const uint64_t c = st->tile_cnt[P + tile];
        if (c == 0) continue;  // uniform across the CTA
        const uint64_t base = st->tile_base[tile];
        const uint64_t start = tile * RESAMPLE_TILE;
        // coalesced loads into shared memory
        for (uint32_t k = tid; k < RESAMPLE_TILE; k += TPB) {
            const uint64_t i = start + k;
            s_cdf[k] = i < n ? offspring[i] : 0;
            s_hap[k] = i < n ? hap_id[i] : 0;
        }
        __syncthreads();
        // inclusive scan: each thread owns PER consecutive items
        uint64_t run = 0;
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) { run += s_cdf[tid * PER + k]; s_cdf[tid * PER + k] = run; }
        uint64_t inc = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        uint64_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < TPB / 32; w++) {
            uint64_t t = s_w[w];
            if (w < (int)wid) woff += t;
            total += t;
        }
        const uint64_t excl = woff + inc - run;
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) s_cdf[tid * PER + k] += excl;
        __syncthreads();
        for (uint64_t j = tid; j < c; j += TPB) {
            Philox g(st->seed, st->compartment, st->generation, PH_RESAMPLE, (base + j) | ((uint64_t)stream_salt << 48));
            const uint64_t r = g.below(total);
            uint32_t lo = 0, hi = RESAMPLE_TILE;  // first index with cdf > r
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (s_cdf[mid] <= r) lo = mid + 1; else hi = mid;
            }
            if (mode == 0) dst[base + j] = s_hap[lo];
            else dst64[base + j] = start + lo;
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
}

// ------------------------------------------------------------------------------------------ garbage collection
// Flattened records have no ancestors to keep alive: a record is live iff an infectant (or a detached
// population) references it.  Compaction runs only when the table or the arena passes its threshold.
__global__ void k_gc_begin(DevState *st, int force) {
    st->gc_active = (force || st->n_hap > st->gc_hap_threshold || st->arena_top > st->gc_arena_threshold) ? 1u : 0u;
}
__global__ void __launch_bounds__(TPB) k_gc_clear(DevState *st) {
    if (!st->gc_active) return;
    const uint64_t nh = st->n_hap;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += (uint64_t)gridDim.x * blockDim.x) st->gc_mark[h] = (h == 0) ? 1u : 0u;
}
__global__ void __launch_bounds__(TPB) k_gc_mark(DevState *st, const uint32_t *__restrict__ ids, uint64_t m, int use_population) {
    if (!st->gc_active) return;
    if (use_population) { ids = st->hap_id; m = st->n; }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) st->gc_mark[ids[i]] = 1u;
}
__global__ void __launch_bounds__(TPB) k_gc_lens(DevState *st) {
    if (!st->gc_active) return;
    const uint64_t nh = st->n_hap;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += (uint64_t)gridDim.x * blockDim.x) st->gc_len[h] = st->gc_mark[h] ? st->h_len[h] : 0u;
}
__global__ void __launch_bounds__(TPB) k_gc_copy(DevState *st) {
    if (!st->gc_active) return;
    const uint64_t nh = st->n_hap;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += (uint64_t)gridDim.x * blockDim.x) {
        if (!st->gc_mark[h]) continue;
        const uint32_t nid = st->gc_newid[h];
        const uint64_t noff = st->gc_newoff[h];
        const uint32_t len = st->h_len[h];
        const uint32_t *src = st->arena + st->h_off[h];
        uint32_t *dst = st->arena_alt + noff;
        for (uint32_t k = 0; k < len; k++) dst[k] = src[k];
        st->h_off_alt[nid] = noff;
        st->h_len_alt[nid] = len;
        const uint32_t par = st->h_parent[h];
        st->h_parent_alt[nid] = (par != NONE32 && st->gc_mark[par]) ? st->gc_newid[par] : NONE32;
        for (uint32_t p = 0; p < st->n_providers; p++) st->h_raw_alt[p][nid] = st->h_raw[p][h];
    }
}
__global__ void __launch_bounds__(TPB) k_gc_remap(DevState *st, uint32_t *ids, uint64_t m, int use_population) {
    if (!st->gc_active) return;
    if (use_population) { ids = st->hap_id; m = st->n; }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) ids[i] = st->gc_newid[ids[i]];
}
// gc_totals[0] = live records, gc_totals[1] = live arena words (written by the two scans)
__global__ void k_gc_end(DevState *st, const uint64_t *gc_totals) {
    if (!st->gc_active) return;
    st->n_hap = gc_totals[0];
    st->arena_top = gc_totals[1];
#define VL_SWAP(a, b) { auto t = st->a; st->a = st->b; st->b = t; }
    VL_SWAP(h_off, h_off_alt) VL_SWAP(h_len, h_len_alt) VL_SWAP(h_parent, h_parent_alt) VL_SWAP(arena, arena_alt)
    for (uint32_t p = 0; p < st->n_providers; p++) VL_SWAP(h_raw[p], h_raw_alt[p])
#undef VL_SWAP
    st->gc_runs += 1;
    st->gc_active = 0;
}

// ------------------------------------------------------------------------------------------ small helpers
__global__ void __launch_bounds__(TPB) k_fill_u32(uint32_t *p, uint64_t n, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void __launch_bounds__(TPB) k_u64_to_u32(DevState *st, const uint64_t *__restrict__ src, uint32_t *__restrict__ dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = src[i];
        if (v > 0xFFFFFFFFull) { raise(st, DE_OFFSPRING_RANGE); v = 0; }
        dst[i] = (uint32_t)v;
    }
}
__global__ void __launch_bounds__(TPB) k_u32_to_u64(const uint32_t *__restrict__ src, uint64_t *__restrict__ dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// per-infectant replication terms for inspection: mapped fitness, host sum and lambda (NaN where uninfected)
__global__ void __launch_bounds__(TPB) k_replication_terms(DevState *st, double *hs_out, double *lam_out) {
    const uint64_t n = st->n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t h = st->host_id[i];
        double s = __longlong_as_double(0x7FF8000000000000ll), l = s;
        if (h != NONE32) { s = st->hsum[h]; l = st->fit[i] * st->host_r0 / s; }
        hs_out[i] = s;
        lam_out[i] = l;
    }
}
__global__ void __launch_bounds__(TPB) k_gather_raw(DevState *st, uint32_t provider, double *out) {
    const uint64_t n = st->n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = st->h_raw[provider][st->hap_id[i]];
}

// sampler test hooks (distribution tests of the device samplers)
__global__ void k_test_poisson(uint64_t seed, double lambda, uint64_t n, uint64_t *out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Philox g(seed, 0, 0, PH_TEST, i);
        out[i] = poisson(g, lambda);
    }
}
__global__ void k_test_binomial(uint64_t seed, uint64_t nn, double p, uint64_t n, uint64_t *out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Philox g(seed, 0, 0, PH_TEST, i);
        out[i] = binomial(g, nn, p);
    }
}
__global__ void k_test_below(uint64_t seed, uint64_t bound, uint64_t n, uint64_t *out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Philox g(seed, 0, 0, PH_TEST, i);
        out[i] = g.below(bound);
    }
}

}  // namespace vl
