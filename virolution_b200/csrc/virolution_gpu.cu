// virolution_gpu.cu — the C ABI of include/virolution_gpu.h on top of the sm_100a kernels.
//
// One vl_sim = one compartment (BasicSimulation, reference src/simulation.rs:221-256) resident on one GPU:
// it owns a stream, the structure-of-arrays population, the host map, the haplotype table/arena and the
// scratch of every kernel.  All sizes that change from generation to generation (population size, number
// of haplotypes, arena top) live in DevState on the device, so vl_next_generation / vl_run enqueue whole
// generations without a host round trip; only the trait-level calls that hand data to the host synchronise.
// There is no CPU fallback anywhere in this file: without a device every entry point fails with VL_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "vl_kernels.cuh"
#include "vl_migrate.cuh"
#include "vl_exchange.cuh"
#include "vl_fused.cuh"

using namespace vl;

#define VL_API extern "C" __attribute__((visibility("default")))

struct vl_pop {
    vl_sim *owner;
    uint32_t *ids;   // device, haplotype ids local to owner
    uint64_t n;      // valid after finalisation
    uint64_t cap;
    uint64_t cap_alloc;  // allocated entries of ids (>= cap when the buffer came from the pool)
    uint8_t *image;  // packed image (vl_pop_pack_device), device
    uint64_t image_bytes;
    int pending_slot;  // pinned slot that receives n, -1 when final
};

struct vl_sim {
    int device = 0;
    int n_sms = 148;
    int flags = 0;
    // one fused generation captured as a CUDA graph: small populations are launch-bound (a dozen kernels of a few
    // microseconds each), replaying the graph halves the time per generation there
    cudaGraphExec_t gen_graph = nullptr;
    uint64_t graph_key[5] = {0, 0, 0, 0, 0};
    uint64_t graph_launches = 0;
    bool graph_disabled = false;
    int experiment = 0;  // VL_EXPERIMENT bit mask, A/B timing aid (default 0): 1 = infect before mutate in the fused generation,
                         // 2 = 512-thread bucketed infect, 4 = unstaged bucket sort, 64 = no CUDA-graph replay
    cudaStream_t stream = nullptr;
    std::recursive_mutex mu;
    DevState h;            // host mirror: configuration + buffer pointers (dynamic fields live on the device)
    DevState *d = nullptr;
    uint32_t *count = nullptr;   // H+1 histogram of infections
    uint64_t cap_H = 0;
    uint64_t *scan_work = nullptr;
    uint64_t scan_work_words = 0;
    uint64_t *gc_totals = nullptr;  // [0],[1] gc; [2],[3] pack; [4],[5] reserve; [6] export total
    uint64_t *stage64 = nullptr;    // cap_inf u64: offspring/indices staging
    uint64_t *pin = nullptr;        // pinned host scratch, PIN_SLOTS u64
    std::vector<void *> fixed_allocs;
    // replay logs (device copies), consumed by the next matching phase
    int64_t *rp_inf = nullptr; bool has_rp_inf = false;
    uint64_t *rp_mut_off = nullptr; uint32_t *rp_mut_sites = nullptr; uint8_t *rp_mut_bases = nullptr; bool has_rp_mut = false;
    uint64_t *rp_rec_host = nullptr; uint32_t *rp_rec_i = nullptr, *rp_rec_j = nullptr, *rp_rec_s0 = nullptr, *rp_rec_s1 = nullptr;
    uint8_t *rp_rec_which = nullptr; uint64_t rp_rec_n = 0; bool has_rp_rec = false;
    uint64_t *rp_off = nullptr; bool has_rp_off = false;
    std::vector<vl_pop *> pops;
    struct IdBuf { uint32_t *p; uint64_t cap; };
    std::vector<IdBuf> id_pool;  // id buffers of dropped populations: reused in stream order, no malloc/free per generation
    struct DevBuf { void *p = nullptr; uint64_t bytes = 0; };
    DevBuf mig[9];               // migrant image staging (export: 0-4, import: 5-8), grown on demand
    uint64_t launches = 0;
    uint64_t n_known = 0; bool n_valid = false;   // host knowledge of the population size
    uint64_t n_bound = 0;                         // upper bound of the population size (launch geometry)
    bool csr_valid = false;       // rec/offsets hold the host map of the current population
    bool slices_ordered = false;  // ... and every slice is in the reference's descending order
    bool layout_valid = false;    // rec/offspring/tile_sum describe an offspring map (host map or identity layout)
    bool layout_identity = false; // ... in infectant order (record p = infectant p), what every resample draws from
    bool mutcounts_ready = false; // the fused infect pass already filled the mutation worklist
    bool bucket_pending = false;  // records are bucket-partitioned, the host map is not built yet
    bool uniform_infection = false;  // the host map came from uniform Philox draws over [0, H) (no replay, DevState::fast_infect)
    bool pop_fit_valid = false;   // pop_fit[] holds the mapped reproductive fitness of every infectant of the current population
    bool has_epistasis = false;   // some provider is epistatic (k_mutate<true>)
    bool prefilled = false;       // the last resample drew the next generation's hosts and mutation counts (k_resample_pos<true>)
    bool begun_ahead = false;     // ... and has begun that generation (counter, counters) as well: no k_begin_generation_fused launch
    bool prefill_patch = false;   // ... and records were created after it (imported migrants): k_mutate patches the population itself
    bool prefill_own_done = false; // ... and the items [0, mut_own) of the prepared worklist have been mutated already (k_mutate range 1)
    bool hsum_dirty = false;      // ... and that work was dropped: hsum[parity] holds stale sums, cleared before the next infect pass
    uint32_t plan_smem_tiles = 0;
    uint32_t salt = 0;        // distinguishes the sampling streams of one generation
    uint32_t gens_since_gc = 0, gc_period = 1;
    int next_slot = 128;  // pinned slots [128, PIN_SLOTS) carry detached-population sizes; lower slots are scratch
    // measurement hooks: per-launch CUDA events on this handle's stream
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
    char err[512] = {0};
};

// whatever changes the population, the parameters or the record ids between two generations of the device-resident
// loop drops the infection the last resample prepared (hosts, mutation worklist, per-host sums of the next generation)
static inline void drop_prefill(vl_sim *sim) {
    if (sim->prefilled) { sim->prefilled = false; sim->prefill_own_done = false; sim->hsum_dirty = true; }
}

// records the start event of a profiled launch and returns the stop event to record after it
static cudaEvent_t prof_begin(vl_sim *sim, const char *name) {
    if (sim->prof.size() >= (1u << 20)) return nullptr;
    vl_sim::ProfRec r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return nullptr;
    cudaEventRecord(r.a, sim->stream);
    sim->prof.push_back(r);
    return r.b;
}

static thread_local char g_create_err[512];
// Handles are independent, but three things touch the whole context: creating a handle (allocations, function attributes),
// destroying one (cudaFree synchronises the device) and capturing a generation into a CUDA graph (no other thread may
// synchronise the device meanwhile).  They take turns; nothing on the per-generation path takes this lock.
static std::mutex g_context_mu;
constexpr int PIN_SLOTS = 4096;
constexpr int BSORT_SMEM_BYTES = BUCKET_HOSTS * 4 + BSORT_STAGE_RECS * 10;

static int fail(vl_sim *sim, int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(sim ? sim->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(sim, expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) return fail(sim, VL_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)
#define TRY(expr)                  \
    do {                           \
        int r_ = (expr);           \
        if (r_ != VL_OK) return r_; \
    } while (0)
#define LAUNCH(sim, kernel, grid, block, smem, ...)                         \
    do {                                                                    \
        cudaEvent_t pe_ = (sim)->profiling ? prof_begin(sim, #kernel) : nullptr; \
        kernel<<<(grid), (block), (smem), (sim)->stream>>>(__VA_ARGS__);    \
        if (pe_) cudaEventRecord(pe_, (sim)->stream);                       \
        (sim)->launches++;                                                  \
    } while (0)
// k_mutate<EPI>: the instantiation with the pair-table code only for handles that have an epistatic provider
#define LAUNCH_MUTATE(sim, grid, block, smem, ...)                                       \
    do {                                                                                 \
        if ((sim)->has_epistasis) LAUNCH(sim, k_mutate<true>, grid, block, smem, __VA_ARGS__); \
        else LAUNCH(sim, k_mutate<false>, grid, block, smem, __VA_ARGS__);               \
    } while (0)
// device-resident generation: events with more sites than the per-thread shared scratch holds (> 8 in one genome and
// generation) are set aside by the main pass and worked off by a second, tiny launch — the main pass then addresses its
// scratch as shared memory only
#define LAUNCH_MUTATE_FUSED(sim, grid, block, smem, ...)                                 \
    do {                                                                                 \
        if ((sim)->has_epistasis) {                                                      \
            LAUNCH(sim, (k_mutate<true, true, 1>), grid, block, smem, __VA_ARGS__);      \
            LAUNCH(sim, (k_mutate<true, true, 2>), 8, block, smem, __VA_ARGS__);         \
        } else {                                                                         \
            LAUNCH(sim, (k_mutate<false, true, 1>), grid, block, smem, __VA_ARGS__);     \
            LAUNCH(sim, (k_mutate<false, true, 2>), 8, block, smem, __VA_ARGS__);        \
        }                                                                                \
    } while (0)
// launches of a scope go to another stream of the same handle
struct StreamScope {
    vl_sim *sim;
    cudaStream_t saved;
    StreamScope(vl_sim *s, cudaStream_t other) : sim(s), saved(s->stream) { s->stream = other; }
    ~StreamScope() { sim->stream = saved; }
};
#define ENTER(sim)                                                   \
    if (!(sim)) return fail(nullptr, VL_ERR_ARGUMENT, "null handle"); \
    std::lock_guard<std::recursive_mutex> lock_((sim)->mu);           \
    CU(sim, cudaSetDevice((sim)->device))

static inline unsigned grid_for(const vl_sim *sim, uint64_t n, unsigned per_sm = 8) {
    uint64_t g = (n + TPB - 1) / TPB;
    uint64_t cap = (uint64_t)sim->n_sms * per_sm;
    if (g > cap) g = cap;
    return (unsigned)(g ? g : 1);
}
static inline unsigned grid_tiles(const vl_sim *sim, uint64_t n_items, uint64_t tile, unsigned per_sm = 8) {
    uint64_t g = (n_items + tile - 1) / tile;
    uint64_t cap = (uint64_t)sim->n_sms * per_sm;
    if (g > cap) g = cap;
    return (unsigned)(g ? g : 1);
}

template <typename T>
static int dev_alloc(vl_sim *sim, T **p, uint64_t count, bool fixed = true) {
    void *q = nullptr;
    CU(sim, cudaMalloc(&q, std::max<uint64_t>(count, 1) * sizeof(T)));
    *p = (T *)q;
    if (fixed) sim->fixed_allocs.push_back(q);
    return VL_OK;
}
#define FIELD_OFF(f) offsetof(DevState, f)
template <typename T>
static int read_field(vl_sim *sim, size_t off, T *out) {
    CU(sim, cudaMemcpyAsync(sim->pin, (const char *)sim->d + off, sizeof(T), cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    memcpy(out, sim->pin, sizeof(T));
    return VL_OK;
}
template <typename T>
static int write_field(vl_sim *sim, size_t off, const T &v) {
    // small parameter updates: staged through the pinned scratch, ordered on the stream
    CU(sim, cudaStreamSynchronize(sim->stream));
    memcpy(sim->pin, &v, sizeof(T));
    CU(sim, cudaMemcpyAsync((char *)sim->d + off, sim->pin, sizeof(T), cudaMemcpyHostToDevice, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    return VL_OK;
}

// sticky device error word -> status
static int check_device_error(vl_sim *sim) {
    uint32_t e = 0;
    TRY(read_field(sim, FIELD_OFF(error), &e));
    CU(sim, cudaGetLastError());
    if (!e) return VL_OK;
    uint32_t zero = 0;
    write_field(sim, FIELD_OFF(error), zero);
    if (e & (DE_HAP_CAPACITY | DE_ARENA_CAPACITY | DE_INF_CAPACITY))
        return fail(sim, VL_ERR_CAPACITY, "device capacity exhausted (bits 0x%x: 1=haplotype table, 2=arena, 32=population buffer)", e);
    if (e & DE_ALL_ZERO_WEIGHTS)
        return fail(sim, VL_ERR_REFERENCE_PANIC, "all offspring weights are zero: the reference panics in WeightedAliasIndex::new (src/core/population.rs:386)");
    if (e & DE_UNDEFINED_OFFSPRING)
        return fail(sim, VL_ERR_REFERENCE_PANIC, "Poisson::new failed on a non-zero fitness: the reference leaves the f64 bit pattern as offspring count (src/simulation.rs:146-152)");
    if (e & DE_SAME_BASE) return fail(sim, VL_ERR_REFERENCE_PANIC, "assert_ne!(from, to) in create_descendant (src/core/haplotype.rs:250)");
    if (e & DE_OFFSPRING_RANGE) return fail(sim, VL_ERR_IMPLEMENTATION, "offspring count does not fit 32 bits");
    if (e & DE_REPLAY) return fail(sim, VL_ERR_ARGUMENT, "malformed replay log or index out of range");
    if (e & DE_PLAN_TIMEOUT) return fail(sim, VL_ERR_CUDA, "the CTAs of the resample plan did not become resident together within 2 s (grid barrier)");
    if (e & DE_EXCHANGE_TIMEOUT) return fail(sim, VL_ERR_CUDA, "migrant exchange timed out: a peer's block did not arrive");
    if (e & DE_TILE_OVERFLOW) return fail(sim, VL_ERR_IMPLEMENTATION, "host tile overflow (> 13 sigma under uniform infection): rerun with VL_FLAG_GENERAL_REPLICATE");
    if (e & DE_BUCKET_OVERFLOW) return fail(sim, VL_ERR_IMPLEMENTATION, "host bucket overflow (12 sigma event): rerun with VL_FLAG_NO_BUCKETS");
    return fail(sim, VL_ERR_IMPLEMENTATION, "unknown device error 0x%x", e);
}

// stream-ordered clear of a (4-byte aligned) device range by a kernel
static inline void zero_device(vl_sim *sim, void *p, uint64_t bytes) {
    const uint64_t n = (bytes + 3) / 4;
    const unsigned g = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (n + TPB - 1) / TPB), (uint64_t)sim->n_sms * 8);
    k_zero_words<<<g, TPB, 0, sim->stream>>>((uint32_t *)p, n);
}
static int zero_scan_work(vl_sim *sim, uint64_t n_items_bound) {
    uint64_t words = 2 + (n_items_bound + SCAN_TILE - 1) / SCAN_TILE;
    if (words > sim->scan_work_words) {
        CU(sim, cudaStreamSynchronize(sim->stream));
        if (sim->scan_work) cudaFree(sim->scan_work);
        sim->scan_work = nullptr;
        sim->scan_work_words = words + words / 2;
        CU(sim, cudaMalloc((void **)&sim->scan_work, sim->scan_work_words * 8));
    }
    zero_device(sim, sim->scan_work, words * 8);
    return VL_OK;
}
template <typename OutT, bool EXCL>
static int launch_scan(vl_sim *sim, const uint32_t *in, OutT *out, const uint64_t *n_ptr, uint64_t n_add, uint64_t bound,
                       uint64_t *total_out, int write_end, const uint32_t *active) {
    TRY(zero_scan_work(sim, bound + 1));
    unsigned g = grid_tiles(sim, bound + 1, SCAN_TILE, 4);
    LAUNCH(sim, (k_scan_u32<OutT, EXCL>), g, SCAN_THREADS, 0, in, out, n_ptr, n_add, sim->scan_work, total_out, write_end, active);
    return VL_OK;
}

// ---------------------------------------------------------------------------------------------- buffers
static void free_population_buffers(vl_sim *sim) {
    DevState &h = sim->h;
    void *olds[] = {h.hap_id, h.hap_id_next, h.host_id, h.rank, h.rec, h.heavy_hosts, h.medium_hosts, h.mut_list, h.big_list, h.lam, h.hs, h.offspring,
                    h.tile_sum, h.tile_tree, h.tile_cnt, h.tile_node, h.tile_base, sim->stage64, h.tmp_rec};
    for (void *p : olds) if (p) cudaFree(p);
    h.tmp_rec = nullptr;
    if (h.pop_fit) cudaFree(h.pop_fit);
    if (h.pop_fit_next) cudaFree(h.pop_fit_next);
    h.pop_fit = h.pop_fit_next = nullptr;
}
// one reproductive provider for every host (or none): the mapped fitness of an infectant does not depend on its host
static bool single_reproductive_provider(const DevState &h) {
    for (uint32_t s = 1; s < h.n_specs; s++) if (h.specs[s].repro != h.specs[0].repro) return false;
    return true;
}
// device-resident generation (vl_fused.cuh): uniform infection, Binomial(L, mu) in the inversion regime, no recombination,
// fitness independent of the host
static bool fused_possible(const vl_sim *sim) {
    const DevState &h = sim->h;
    return h.fast_infect && h.binv.ok && !(h.host_recombination_rate > 0.0) && single_reproductive_provider(h) && h.hsum[0] != nullptr &&
           !(sim->flags & (VL_FLAG_NO_BUCKETS | VL_FLAG_GENERAL_REPLICATE | VL_FLAG_NO_TILES)) && !(sim->experiment & 8);
}
// tiles the plan/resample arrays can describe: position tiles of the largest population, or host tiles of as many hosts
static inline uint64_t tile_capacity(uint64_t cap_inf) { return (cap_inf + HT_MIN_HOSTS - 1) / HT_MIN_HOSTS + 2; }
static bool buckets_possible(const vl_sim *sim) {
    const DevState &h = sim->h;
    return h.fast_infect && h.H >= 4 * (uint64_t)BUCKET_HOSTS && h.n_buckets <= (uint32_t)MAX_BUCKETS && !(sim->flags & VL_FLAG_NO_BUCKETS);
}
static int alloc_bucket_buffers(vl_sim *sim) {
    DevState &h = sim->h;
    if (h.tmp_rec) cudaFree(h.tmp_rec);
    h.tmp_rec = nullptr;
    h.cap_tmp = 0;
    if (!buckets_possible(sim)) return VL_OK;
    h.cap_tmp = (uint64_t)h.n_buckets * bucket_stride(h.cap_inf, h.H) + 64;
    TRY(dev_alloc(sim, &h.tmp_rec, h.cap_tmp, false));
    return VL_OK;
}
static int alloc_population_buffers(vl_sim *sim, uint64_t cap) {
    DevState &h = sim->h;
    free_population_buffers(sim);
    const uint64_t nt = tile_capacity(cap);
    uint64_t P = 1;
    while (P < nt) P <<= 1;
    TRY(dev_alloc(sim, &h.hap_id, cap + 4, false));
    TRY(dev_alloc(sim, &h.hap_id_next, cap + 4, false));
    TRY(dev_alloc(sim, &h.host_id, cap + 4, false));
    TRY(dev_alloc(sim, &h.rank, cap + 4, false));
    TRY(dev_alloc(sim, &h.rec, cap + 4, false));
    TRY(dev_alloc(sim, &h.heavy_hosts, cap / THREAD_TIER + 1, false));
    TRY(dev_alloc(sim, &h.medium_hosts, cap / 5 + 1, false));
    TRY(dev_alloc(sim, &h.mut_list, cap + 4, false));
    TRY(dev_alloc(sim, &h.big_list, cap + 4, false));
    TRY(dev_alloc(sim, &h.lam, cap + 4, false));
    TRY(dev_alloc(sim, &h.hs, cap + 4, false));
    TRY(dev_alloc(sim, &h.offspring, cap + 4, false));
    TRY(dev_alloc(sim, &h.tile_sum, nt, false));
    TRY(dev_alloc(sim, &h.tile_tree, 2 * P, false));
    TRY(dev_alloc(sim, &h.tile_cnt, nt * MAX_PARTS, false));
    TRY(dev_alloc(sim, &h.tile_node, 2 * P, false));
    TRY(dev_alloc(sim, &h.tile_base, nt * MAX_PARTS, false));
    h.tile_stride = nt;
    TRY(dev_alloc(sim, &sim->stage64, cap + 4, false));
    TRY(dev_alloc(sim, &h.pop_fit, cap + 4, false));
    TRY(dev_alloc(sim, &h.pop_fit_next, cap + 4, false));
    h.cap_inf = cap;
    TRY(alloc_bucket_buffers(sim));
    return VL_OK;
}
// push the host mirror's population pointers/capacity into the device state
static int push_population_pointers(vl_sim *sim) {
    CU(sim, cudaStreamSynchronize(sim->stream));
    DevState tmp;
    CU(sim, cudaMemcpy(&tmp, sim->d, sizeof(DevState), cudaMemcpyDeviceToHost));
    const DevState &h = sim->h;
    tmp.hap_id = h.hap_id; tmp.hap_id_next = h.hap_id_next; tmp.host_id = h.host_id; tmp.rank = h.rank; tmp.rec = h.rec;
    tmp.heavy_hosts = h.heavy_hosts; tmp.medium_hosts = h.medium_hosts; tmp.mut_list = h.mut_list; tmp.big_list = h.big_list; tmp.lam = h.lam; tmp.hs = h.hs; tmp.offspring = h.offspring;
    tmp.tile_sum = h.tile_sum; tmp.tile_tree = h.tile_tree; tmp.tile_cnt = h.tile_cnt; tmp.tile_node = h.tile_node; tmp.tile_base = h.tile_base;
    tmp.tile_stride = h.tile_stride;
    tmp.cap_inf = h.cap_inf; tmp.offsets = h.offsets; tmp.H = h.H; tmp.fast_infect = h.fast_infect;
    tmp.n_buckets = h.n_buckets; tmp.tmp_rec = h.tmp_rec; tmp.cap_tmp = h.cap_tmp;
    tmp.pop_fit = h.pop_fit; tmp.pop_fit_next = h.pop_fit_next; tmp.hsum[0] = h.hsum[0]; tmp.hsum[1] = h.hsum[1];
    CU(sim, cudaMemcpy(sim->d, &tmp, sizeof(DevState), cudaMemcpyHostToDevice));
    return VL_OK;
}
static int alloc_host_buffers(vl_sim *sim, uint64_t H) {
    DevState &h = sim->h;
    if (h.offsets) cudaFree(h.offsets);
    if (sim->count) cudaFree(sim->count);
    TRY(dev_alloc(sim, &h.offsets, H + 8, false));
    TRY(dev_alloc(sim, &sim->count, H + 8, false));
    // per-host fitness sums of the device-resident generation (vl_fused.cuh), both cleared: a generation accumulates into
    // hsum[parity] and clears the other.  (Pinning the array in L2 with an access-policy window was measured and dropped:
    // the 83 MB set aside for it cost the other kernels more than it saved — 1.66 against 0.96 ms per generation at K2.)
    if (h.hsum[0]) cudaFree(h.hsum[0]);
    h.hsum[0] = h.hsum[1] = nullptr;
    if (!(sim->flags & VL_FLAG_NO_TILES)) {
        const uint64_t per = (H + 8 + 15) & ~15ull;
        TRY(dev_alloc(sim, &h.hsum[0], 2 * per, false));
        h.hsum[1] = h.hsum[0] + per;
        CU(sim, cudaMemset(h.hsum[0], 0, 2 * per * sizeof(double)));
    }
    sim->cap_H = H;
    return VL_OK;
}
// every host belongs to a spec, no spec has an infective provider and n_hits == 1: each draw infects its host
// (src/simulation.rs:56-65 degenerates to one Uniform(0, H) draw), so hosts are uniform over [0, H)
static void update_fast_infect(vl_sim *sim) {
    DevState &h = sim->h;
    bool fast = h.n_hits == 1 && h.H > 0;
    uint64_t covered = 0;  // specs in ascending order must tile [0, H)
    for (uint32_t s = 0; s < h.n_specs && fast; s++) {
        if (h.specs[s].infect >= 0) fast = false;
        if (h.specs[s].lo > covered) fast = false;
        covered = std::max<uint64_t>(covered, h.specs[s].hi);
    }
    if (covered < h.H) fast = false;
    h.fast_infect = fast ? 1u : 0u;
    h.n_buckets = (uint32_t)std::min<uint64_t>((h.H + BUCKET_HOSTS - 1) >> BUCKET_SHIFT, 0xFFFFFFFFull);
}

static int add_provider(vl_sim *sim, const vl_fitness *f, uint64_t L, int *index_out) {
    *index_out = -1;
    if (f->kind == VL_FITNESS_NONE) return VL_OK;
    if (f->kind < 0 || f->kind > VL_FITNESS_EPISTATIC) return fail(sim, VL_ERR_ARGUMENT, "unknown fitness kind %d", f->kind);
    DevState &h = sim->h;
    if (h.n_providers >= (uint32_t)MAX_PROVIDERS) return fail(sim, VL_ERR_ARGUMENT, "too many fitness providers");
    DevProvider &p = h.prov[h.n_providers];
    memset(&p, 0, sizeof(p));
    p.kind = f->kind;
    p.util_kind = f->utility.kind;
    p.upper = f->utility.upper;
    p.hill_n = f->utility.n;
    p.hill_k = (1.0 - f->utility.p) / f->utility.p;  // UtilityFunction::new_hill, src/core/fitness/utility.rs:26-45
    {
        // UtilityFunction::apply, Algebraic arm (utility.rs:51-54): `let factor = 1. / (upper - 1.)` then
        // `upper * factor * f / ...` — both constants are formed here with the same two roundings
        volatile double factor = 1. / (f->utility.upper - 1.);
        volatile double uf = f->utility.upper * factor;
        p.alg_factor = factor;
        p.alg_uf = uf;
    }
    if (f->kind >= VL_FITNESS_NONEPISTATIC) {
        if (!f->table) return fail(sim, VL_ERR_INITIALIZATION, "fitness table missing");
        double *t = nullptr;
        TRY(dev_alloc(sim, &t, L * 4));
        CU(sim, cudaMemcpy(t, f->table, L * 4 * sizeof(double), cudaMemcpyHostToDevice));
        p.table = t;
    }
    if (f->kind == VL_FITNESS_EPISTATIC) {
        // EpistasisTable::from_vec (src/core/fitness/epistasis.rs:44-58): symmetric nested map, a later entry
        // for the same ordered key pair overwrites an earlier one
        struct Dir { uint32_t k1, k2; double v; uint64_t seq; };
        std::vector<Dir> d;
        d.reserve(2 * f->n_epi);
        for (uint64_t i = 0; i < f->n_epi; i++) {
            const vl_epi_entry &e = f->epi[i];
            if (e.pos1 >= L || e.pos2 >= L || e.base1 > 3 || e.base2 > 3) return fail(sim, VL_ERR_INITIALIZATION, "epistasis entry %llu out of range", (unsigned long long)i);
            uint32_t a = ekey((uint32_t)e.pos1, e.base1), b = ekey((uint32_t)e.pos2, e.base2);
            d.push_back({a, b, e.value, d.size()});
            d.push_back({b, a, e.value, d.size()});
        }
        std::sort(d.begin(), d.end(), [](const Dir &x, const Dir &y) {
            if (x.k1 != y.k1) return x.k1 < y.k1;
            if (x.k2 != y.k2) return x.k2 < y.k2;
            return x.seq < y.seq;
        });
        std::vector<EpiDir> out;
        for (size_t i = 0; i < d.size(); i++) {
            if (i + 1 < d.size() && d[i + 1].k1 == d[i].k1 && d[i + 1].k2 == d[i].k2) continue;
            out.push_back({d[i].k1, d[i].k2, d[i].v});
        }
        EpiDir *dev = nullptr;
        TRY(dev_alloc(sim, &dev, out.size()));
        if (!out.empty()) CU(sim, cudaMemcpy(dev, out.data(), out.size() * sizeof(EpiDir), cudaMemcpyHostToDevice));
        p.epi = dev;
        p.n_epi = (uint32_t)out.size();
        // row index: first entry of every (position, base) key
        std::vector<uint32_t> start((size_t)L * 4 + 1, 0);
        for (const EpiDir &e : out) start[e.k1 + 1]++;
        for (size_t k = 0; k < (size_t)L * 4; k++) start[k + 1] += start[k];
        uint32_t *dstart = nullptr;
        TRY(dev_alloc(sim, &dstart, start.size()));
        CU(sim, cudaMemcpy(dstart, start.data(), start.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        p.epi_start = dstart;
        p.n_keys = (uint32_t)(L * 4);
        sim->has_epistasis = true;
    }
    *index_out = (int)h.n_providers++;
    return VL_OK;
}

static void destroy_sim(vl_sim *sim) {
    if (!sim) return;
    cudaSetDevice(sim->device);
    if (sim->stream) cudaStreamSynchronize(sim->stream);
    for (vl_pop *p : sim->pops) {
        if (p->ids) cudaFree(p->ids);
        if (p->image) cudaFree(p->image);
        delete p;
    }
    if (sim->gen_graph) cudaGraphExecDestroy(sim->gen_graph);
    for (auto &b : sim->id_pool) cudaFree(b.p);
    free_population_buffers(sim);
    if (sim->h.hsum[0]) cudaFree(sim->h.hsum[0]);
    if (sim->h.offsets) cudaFree(sim->h.offsets);
    if (sim->count) cudaFree(sim->count);
    for (void *p : sim->fixed_allocs) cudaFree(p);
    if (sim->scan_work) cudaFree(sim->scan_work);
    void *rp[] = {sim->rp_inf, sim->rp_mut_off, sim->rp_mut_sites, sim->rp_mut_bases, sim->rp_rec_host, sim->rp_rec_i, sim->rp_rec_j,
                  sim->rp_rec_s0, sim->rp_rec_s1, sim->rp_rec_which, sim->rp_off};
    for (void *p : rp) if (p) cudaFree(p);
    for (auto &b : sim->mig) if (b.p) cudaFree(b.p);
    if (sim->pin) cudaFreeHost(sim->pin);
    for (auto &r : sim->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (sim->timer_a) { cudaEventDestroy(sim->timer_a); cudaEventDestroy(sim->timer_b); }
    if (sim->stream) cudaStreamDestroy(sim->stream);
    delete sim;
}

// ---------------------------------------------------------------------------------------------- lifecycle
VL_API int vl_abi_version(void) { return 1; }

VL_API const char *vl_last_error(const vl_sim *sim) { return sim ? sim->err : g_create_err; }

VL_API int vl_create(const vl_config *cfg, vl_sim **out) {
    if (!cfg || !out) return fail(nullptr, VL_ERR_ARGUMENT, "null argument");
    *out = nullptr;
    const uint64_t L = cfg->n_sites;
    if (L == 0 || L >= (1ull << 24)) return fail(nullptr, VL_ERR_INITIALIZATION, "n_sites must be in [1, 2^24)");
    if (!cfg->wildtype) return fail(nullptr, VL_ERR_INITIALIZATION, "wildtype missing");
    if (cfg->n_specs == 0 || cfg->n_specs > (uint32_t)MAX_SPECS) return fail(nullptr, VL_ERR_INITIALIZATION, "n_specs must be in [1, %d]", MAX_SPECS);
    const vl_parameters &par = cfg->parameters;
    if (par.host_population_size >= 0xFFFFFFFFull) return fail(nullptr, VL_ERR_INITIALIZATION, "host_population_size must fit 32 bits");
    for (uint64_t i = 0; i < L; i++) if (cfg->wildtype[i] > 3) return fail(nullptr, VL_ERR_INITIALIZATION, "wildtype symbol %llu out of range", (unsigned long long)i);
    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev == 0) return fail(nullptr, VL_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, VL_ERR_ARGUMENT, "device ordinal %d out of range", cfg->device);
    std::lock_guard<std::mutex> context_lock(g_context_mu);
    vl_sim *sim = new vl_sim();
    memset(&sim->h, 0, sizeof(DevState));
    sim->device = cfg->device;
    sim->flags = cfg->flags;
    auto bail = [&](int code) { snprintf(g_create_err, sizeof g_create_err, "%s", sim->err); destroy_sim(sim); return code; };
#define CTRY(expr) do { int r__ = (expr); if (r__ != VL_OK) return bail(r__); } while (0)
#define CCU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { fail(sim, VL_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); return bail(VL_ERR_CUDA); } } while (0)
    CCU(cudaSetDevice(sim->device));
    CCU(cudaDeviceGetAttribute(&sim->n_sms, cudaDevAttrMultiProcessorCount, sim->device));
    CCU(cudaStreamCreateWithFlags(&sim->stream, cudaStreamNonBlocking));
    CCU(cudaMallocHost((void **)&sim->pin, PIN_SLOTS * sizeof(uint64_t)));
    DevState &h = sim->h;
    h.L = (uint32_t)L;
    h.n_specs = cfg->n_specs;
    h.compartment = cfg->compartment_id;
    h.seed = cfg->seed;
    h.host_mutation_rate = par.mutation_rate;
    h.host_recombination_rate = par.recombination_rate;
    h.host_r0 = par.basic_reproductive_number;
    memcpy(h.host_subst, par.substitution_matrix, sizeof h.host_subst);
    h.H = par.host_population_size;
    h.n_hits = par.n_hits;
    h.dilution = par.dilution;
    h.max_population = par.max_population;
    {
        uint8_t *wt = nullptr;
        CTRY(dev_alloc(sim, &wt, L));
        CCU(cudaMemcpy(wt, cfg->wildtype, L, cudaMemcpyHostToDevice));
        h.wildtype = wt;
    }
    for (uint32_t s = 0; s < cfg->n_specs; s++) {
        h.specs[s].lo = cfg->specs[s].lo;
        h.specs[s].hi = cfg->specs[s].hi;
        CTRY(add_provider(sim, &cfg->specs[s].reproductive, L, &h.specs[s].repro));
        CTRY(add_provider(sim, &cfg->specs[s].infective, L, &h.specs[s].infect));
    }
    // ---- capacities
    uint64_t cap_inf = cfg->capacity_infectants ? cfg->capacity_infectants : std::max<uint64_t>(2 * par.max_population, 4096);
    // room for several generations of new records between compactions (HBM is plentiful: 180 GB)
    // The record tables get up to ~45 % of the free device memory (≈ 340 bytes per record with both copies and its share of
    // the arena), between 6 and 16 records of room per infectant slot: a compaction costs a pass over the whole table plus
    // a copy of the live records, so the rarer it runs the cheaper a generation is.
    uint64_t hap_room = 6;
    {
        size_t free_b = 0, total_b = 0;
        CCU(cudaMemGetInfo(&free_b, &total_b));
        const double fit = 0.45 * (double)free_b / ((double)cap_inf * 230.0);
        hap_room = (uint64_t)std::min(16.0, std::max(6.0, floor(fit)));
    }
    uint64_t cap_hap = cfg->capacity_haplotypes ? cfg->capacity_haplotypes : hap_room * cap_inf + 65536;
    if (cap_hap >= 0xFFFFFFF0ull) cap_hap = 0xFFFFFFF0ull;
    uint64_t cap_arena = cfg->capacity_arena_words ? cfg->capacity_arena_words : std::max<uint64_t>(16 * cap_hap, 1ull << 20);
    {
        size_t free_b = 0, total_b = 0;
        CCU(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t P = std::max<uint32_t>(h.n_providers, 1);
        auto need = [&](uint64_t ch, uint64_t ca) { return cap_inf * 100 + ch * (130 + 32 * P) + ca * 8; };
        if (!cfg->capacity_haplotypes || !cfg->capacity_arena_words) {
            for (int it = 0; it < 64 && need(cap_hap, cap_arena) > (uint64_t)(0.7 * (double)free_b); it++) {
                if (!cfg->capacity_arena_words && cap_arena > (1ull << 20)) cap_arena = cap_arena * 3 / 4;
                if (!cfg->capacity_haplotypes && cap_hap > cap_inf + 65536) cap_hap = std::max(cap_inf + 65536, cap_hap * 3 / 4);
            }
        }
        if (need(cap_hap, cap_arena) > (uint64_t)(0.95 * (double)free_b)) {
            fail(sim, VL_ERR_CAPACITY, "configuration needs %.1f GB of device memory, %.1f GB free", need(cap_hap, cap_arena) / 1e9, free_b / 1e9);
            return bail(VL_ERR_CAPACITY);
        }
    }
    h.cap_hap = cap_hap;
    h.cap_arena = cap_arena;
    update_fast_infect(sim);
    h.binv = make_binv(L, par.mutation_rate);
    CCU(cudaFuncSetAttribute(k_plan_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CCU(cudaFuncSetAttribute(k_host_sum_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SORT_MAX * 16));
    CCU(cudaFuncSetAttribute(k_mutate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mutate_smem_bytes(MUT_WT_MAX_SITES)));
    CCU(cudaFuncSetAttribute(k_mutate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mutate_smem_bytes(MUT_WT_MAX_SITES)));
    CCU(cudaFuncSetAttribute((k_mutate<false, true, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mutate_smem_bytes(MUT_WT_MAX_SITES)));
    CCU(cudaFuncSetAttribute((k_mutate<true, true, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mutate_smem_bytes(MUT_WT_MAX_SITES)));
    CCU(cudaFuncSetAttribute((k_mutate<false, true, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mutate_smem_bytes(MUT_WT_MAX_SITES)));
    CCU(cudaFuncSetAttribute((k_mutate<true, true, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mutate_smem_bytes(MUT_WT_MAX_SITES)));
    CCU(cudaFuncSetAttribute((k_resample_pos<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CL_CAP * sizeof(uint4))));
    CCU(cudaFuncSetAttribute((k_resample_pos<true, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CL_CAP * sizeof(uint4))));
    CCU(cudaFuncSetAttribute((k_resample_pos<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CL_CAP * sizeof(uint4))));
    CCU(cudaFuncSetAttribute((k_resample_pos<true, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CL_CAP * sizeof(uint4))));
    CCU(cudaFuncSetAttribute(k_ex_export, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EXP_SMEM_BYTES));
    sim->plan_smem_tiles = 200 * 1024 / 8;
    h.gc_hap_threshold = cap_hap / 2;
    h.gc_arena_threshold = cap_arena / 2;
    {   // compaction is cheap (bit map + ranks, cost ~ population + live records) and a compact table keeps the random
        // reads of k_mutate closer together: run it every few generations' worth of new records, not when memory runs out
        const char *e = getenv("VL_GC_RECORDS");
        const uint64_t target = e ? (uint64_t)atof(e) : 2 * cap_inf;
        if (!cfg->capacity_haplotypes && target > 0) h.gc_hap_threshold = std::min<uint64_t>(h.gc_hap_threshold, std::max<uint64_t>(target, 1u << 16));
    }
    CTRY(dev_alloc(sim, &h.bucket_fill, MAX_BUCKETS + 1));
    CTRY(dev_alloc(sim, &h.bucket_pos, MAX_BUCKETS + 2));
    CCU(cudaFuncSetAttribute((k_bucket_sort<1024, BSORT_STAGE_RECS>), cudaFuncAttributeMaxDynamicSharedMemorySize, BSORT_SMEM_BYTES));
    { const char *e = getenv("VL_EXPERIMENT"); sim->experiment = e ? atoi(e) : 0; }
    CTRY(alloc_population_buffers(sim, cap_inf));
    CTRY(alloc_host_buffers(sim, h.H));
    CTRY(dev_alloc(sim, &h.head, cap_hap)); CTRY(dev_alloc(sim, &h.head_alt, cap_hap));
    for (uint32_t p = 0; p < h.n_providers; p++) {
        CTRY(dev_alloc(sim, &h.h_raw[p], cap_hap)); CTRY(dev_alloc(sim, &h.h_raw_alt[p], cap_hap));
        CTRY(dev_alloc(sim, &h.h_fit[p], cap_hap)); CTRY(dev_alloc(sim, &h.h_fit_alt[p], cap_hap));
    }
    CTRY(dev_alloc(sim, &h.arena, cap_arena + 32)); CTRY(dev_alloc(sim, &h.arena_alt, cap_arena + 32));  // (padded: the search window of entries_lower_interp may overhang a list)
    {
        const uint64_t gw = cap_hap / 32 + 8;  // one bit-map word per 32 records
        CTRY(dev_alloc(sim, &h.gc_bits, gw)); CTRY(dev_alloc(sim, &h.gc_wcnt, gw)); CTRY(dev_alloc(sim, &h.gc_wlen, gw));
        CTRY(dev_alloc(sim, &h.gc_cbase, gw)); CTRY(dev_alloc(sim, &h.gc_lbase, gw));
    }
    CTRY(dev_alloc(sim, &sim->gc_totals, 16));
    CCU(cudaMemset(sim->gc_totals, 0, 16 * 8));
    sim->scan_work_words = 8 + (std::max<uint64_t>(std::max<uint64_t>(cap_hap / 32 + 8, h.H), cap_inf) + SCAN_TILE) / SCAN_TILE;
    CTRY(dev_alloc(sim, &sim->scan_work, sim->scan_work_words, false));
    // haplotype 0 = wildtype: no entries, raw fitness = empty product (src/core/fitness/table.rs:72-79)
    h.n_hap = 1;
    h.arena_top = 0;
    CTRY(dev_alloc(sim, &sim->d, 1));
    CCU(cudaMemcpy(sim->d, &h, sizeof(DevState), cudaMemcpyHostToDevice));
    k_init_wildtype<<<1, 1, 0, sim->stream>>>(sim->d);  // raw fitness 1 and its mapped value under every provider
    // garbage-collection cadence: expected new records per generation from p_mut = 1-(1-mu)^L, with slack
    {
        double p_mut = 1.0 - pow(1.0 - par.mutation_rate, (double)L);
        double per_gen = (double)cap_inf * std::min(1.0, 3.0 * p_mut + 0.01 + (par.recombination_rate > 0 ? 0.05 : 0.0));
        double room = (double)std::min<uint64_t>(cap_hap - h.gc_hap_threshold, h.gc_hap_threshold);
        double period = floor(room / std::max(per_gen, 1.0));
        sim->gc_period = (uint32_t)std::min(64.0, std::max(1.0, period));
    }
    sim->n_known = 0;
    sim->n_valid = true;
    CCU(cudaStreamSynchronize(sim->stream));  // (not the device: another handle's stream may be capturing a graph)
#undef CTRY
#undef CCU
    *out = sim;
    return VL_OK;
}

VL_API int vl_destroy(vl_sim *sim) {
    if (!sim) return VL_OK;
    std::lock_guard<std::mutex> context_lock(g_context_mu);
    destroy_sim(sim);
    return VL_OK;
}

VL_API int vl_synchronize(vl_sim *sim) {
    ENTER(sim);
    CU(sim, cudaStreamSynchronize(sim->stream));
    return check_device_error(sim);
}

// ---------------------------------------------------------------------------------------------- trait: scalars
VL_API int vl_increment_generation(vl_sim *sim) {
    ENTER(sim);
    drop_prefill(sim);
    LAUNCH(sim, k_tick, 1, 1, 0, sim->d);
    sim->salt = 0;
    return VL_OK;
}
VL_API int vl_get_generation(const vl_sim *csim, uint64_t *out) {
    vl_sim *sim = const_cast<vl_sim *>(csim);
    ENTER(sim);
    uint32_t g = 0;
    TRY(read_field(sim, FIELD_OFF(generation), &g));
    *out = g;
    return VL_OK;
}
VL_API int vl_set_generation(vl_sim *sim, uint64_t generation) {
    ENTER(sim);
    drop_prefill(sim);
    return write_field(sim, FIELD_OFF(generation), (uint32_t)generation);
}
VL_API int vl_set_parameters(vl_sim *sim, const vl_parameters *p) {
    ENTER(sim);
    if (!p) return fail(sim, VL_ERR_ARGUMENT, "null parameters");
    drop_prefill(sim);
    if (p->host_population_size >= 0xFFFFFFFFull) return fail(sim, VL_ERR_ARGUMENT, "host_population_size must fit 32 bits");
    // BasicSimulation::set_parameters (src/simulation.rs:355-363) replaces the simulation's copy only; the
    // BasicHost copies (mutation/recombination rate, R0, substitution matrix) keep their construction values.
    DevState &h = sim->h;
    CU(sim, cudaStreamSynchronize(sim->stream));
    if (p->host_population_size > sim->cap_H) TRY(alloc_host_buffers(sim, p->host_population_size));
    h.H = p->host_population_size;
    TRY(push_population_pointers(sim));
    TRY(write_field(sim, FIELD_OFF(n_hits), (uint64_t)p->n_hits));
    TRY(write_field(sim, FIELD_OFF(dilution), p->dilution));
    TRY(write_field(sim, FIELD_OFF(max_population), (uint64_t)p->max_population));
    h.n_hits = p->n_hits; h.dilution = p->dilution; h.max_population = p->max_population;
    update_fast_infect(sim);
    TRY(alloc_bucket_buffers(sim));
    TRY(push_population_pointers(sim));
    sim->csr_valid = false;
    sim->bucket_pending = false;
    if (sim->gen_graph) { cudaGraphExecDestroy(sim->gen_graph); sim->gen_graph = nullptr; }  // captured with the old parameters
    return VL_OK;
}
static int population_len(vl_sim *sim, uint64_t *out) {
    if (!sim->n_valid) {
        TRY(read_field(sim, FIELD_OFF(n), &sim->n_known));
        sim->n_valid = true;
    }
    *out = sim->n_known;
    return VL_OK;
}
VL_API int vl_population_len(vl_sim *sim, uint64_t *out) {
    ENTER(sim);
    TRY(population_len(sim, out));
    return VL_OK;
}
VL_API int vl_set_population_wildtype(vl_sim *sim, uint64_t n) {
    ENTER(sim);
    drop_prefill(sim);
    if (n > sim->h.cap_inf) {
        CU(sim, cudaStreamSynchronize(sim->stream));
        if (!sim->pops.empty()) return fail(sim, VL_ERR_CAPACITY, "population of %llu exceeds capacity_infectants %llu", (unsigned long long)n, (unsigned long long)sim->h.cap_inf);
        TRY(alloc_population_buffers(sim, n + n / 4));
        TRY(push_population_pointers(sim));
    }
    LAUNCH(sim, k_fill_population, grid_for(sim, n), TPB, 0, sim->d, n, 0u);
    sim->pop_fit_valid = false;
    sim->n_known = n;
    sim->n_bound = n;
    sim->n_valid = true;
    sim->csr_valid = false;
    sim->slices_ordered = false;
    sim->layout_valid = false;
    sim->mutcounts_ready = false;
    return VL_OK;
}

// ---------------------------------------------------------------------------------------------- phases
static int zero_field32(vl_sim *sim, size_t off) {
    zero_device(sim, (char *)sim->d + off, 4);
    return VL_OK;
}
// n_medium, n_heavy, fused_failed (adjacent in DevState)
static int zero_replicate_counters(vl_sim *sim) {
    static_assert(FIELD_OFF(n_heavy) == FIELD_OFF(n_medium) + 4 && FIELD_OFF(fused_failed) == FIELD_OFF(n_medium) + 8, "counters must be adjacent");
    zero_device(sim, (char *)sim->d + FIELD_OFF(n_medium), 12);
    return VL_OK;
}
// descending infectant order inside every slice (the reference's back-fill order); part of k_host_fitness in
// the replicate phase, run on its own when recombination or an inspection call needs it earlier
static int ensure_ordered_slices(vl_sim *sim) {
    if (!sim->csr_valid) return fail(sim, VL_ERR_IMPLEMENTATION, "no host map: call vl_infect first");
    if (sim->slices_ordered) return VL_OK;
    TRY(zero_replicate_counters(sim));
    LAUNCH(sim, k_host_sum, grid_for(sim, sim->h.H), TPB, 0, sim->d, 1, 0);
    LAUNCH(sim, k_host_sum_medium, grid_for(sim, sim->n_bound / 5 + 1, 2), TPB, 0, sim->d, 1, 0);
    LAUNCH(sim, k_host_sum_heavy, sim->n_sms, TPB, SMEM_SORT_MAX * 16, sim->d, 1, 0);
    sim->slices_ordered = true;
    return VL_OK;
}
static int enqueue_host_map(vl_sim *sim) {
    // counting sort by host: exclusive scan of the histogram, then every record moves to its CSR position
    const DevState &h = sim->h;
    const unsigned g = grid_for(sim, (sim->n_bound + INFECT_ITEMS - 1) / INFECT_ITEMS);
    TRY((launch_scan<uint32_t, true>(sim, sim->count, h.offsets, (const uint64_t *)((const char *)sim->d + FIELD_OFF(H)), 0, h.H,
                                     (uint64_t *)((char *)sim->d + FIELD_OFF(n_sorted)), 1, nullptr)));
    LAUNCH(sim, k_scatter, g, TPB, 0, sim->d);
    sim->csr_valid = true;
    sim->slices_ordered = false;
    sim->layout_valid = false;
    return VL_OK;
}
// defer_host_map: the fused generation mutates first (the mutation worklist comes out of the infect pass) and
// builds the host map from the already-mutated population, so no record needs patching afterwards
static int enqueue_bucket_host_map(vl_sim *sim) {
    const DevState &h = sim->h;
    LAUNCH(sim, k_bucket_offsets, 1, MAX_BUCKETS / 2, 0, sim->d);
    if (!(sim->experiment & 4)) LAUNCH(sim, (k_bucket_sort<1024, BSORT_STAGE_RECS>), std::min<unsigned>(h.n_buckets, sim->n_sms), 1024, BSORT_SMEM_BYTES, sim->d);
    else LAUNCH(sim, (k_bucket_sort<512, 0>), std::min<unsigned>(h.n_buckets, sim->n_sms * 4), 512, BUCKET_HOSTS * 4, sim->d);
    sim->bucket_pending = false;
    sim->csr_valid = true;
    sim->slices_ordered = false;
    sim->layout_valid = false;
    return VL_OK;
}
static int enqueue_infect(vl_sim *sim, bool defer_host_map = false) {
    const DevState &h = sim->h;
    drop_prefill(sim);
    const bool replay = sim->has_rp_inf || sim->has_rp_mut || sim->has_rp_rec;
    const bool fuse = !replay && h.binv.ok;
    const bool fast = h.fast_infect != 0;
    LAUNCH(sim, k_begin_phase, 1, 1, 0, sim->d);
    sim->bucket_pending = false;
    sim->uniform_infection = fast && !sim->has_rp_inf;
    if (!replay && buckets_possible(sim) && h.cap_tmp >= (uint64_t)h.n_buckets * bucket_stride(sim->n_bound, h.H)) {
        zero_device(sim, h.bucket_fill, (MAX_BUCKETS + 1) * 4);
        const int btpb = (sim->experiment & 2) ? 512 : 1024;
        const unsigned g = (unsigned)std::max<uint64_t>(1, (sim->n_bound + btpb * BUCKET_ITEMS - 1) / (btpb * BUCKET_ITEMS));  // one chunk per CTA
        if (btpb == 1024) {
            if (fuse) LAUNCH(sim, (k_infect_bucket<true, true, 1024>), g, 1024, 0, sim->d);
            else LAUNCH(sim, (k_infect_bucket<false, true, 1024>), g, 1024, 0, sim->d);
        } else {
            if (fuse) LAUNCH(sim, (k_infect_bucket<true, true, 512>), g, 512, 0, sim->d);
            else LAUNCH(sim, (k_infect_bucket<false, true, 512>), g, 512, 0, sim->d);
        }
        sim->mutcounts_ready = fuse;
        sim->csr_valid = false;
        sim->slices_ordered = false;
        sim->layout_valid = false;
        sim->bucket_pending = true;
        if (defer_host_map && fuse && !(h.host_recombination_rate > 0.0)) return VL_OK;
        TRY(enqueue_bucket_host_map(sim));
        if (h.host_recombination_rate > 0.0) TRY(ensure_ordered_slices(sim));
        return VL_OK;
    }
    zero_device(sim, sim->count, (h.H + 2) * 4);
    const unsigned g = grid_for(sim, (sim->n_bound + INFECT_ITEMS - 1) / INFECT_ITEMS);
    const int64_t *rp = sim->has_rp_inf ? sim->rp_inf : nullptr;
    if (fuse && fast) LAUNCH(sim, (k_infect<true, true>), g, TPB, 0, sim->d, sim->count, (const int64_t *)nullptr);
    else if (fuse) LAUNCH(sim, (k_infect<true, false>), g, TPB, 0, sim->d, sim->count, (const int64_t *)nullptr);
    else if (fast) LAUNCH(sim, (k_infect<false, true>), g, TPB, 0, sim->d, sim->count, rp);
    else LAUNCH(sim, (k_infect<false, false>), g, TPB, 0, sim->d, sim->count, rp);
    sim->has_rp_inf = false;
    sim->mutcounts_ready = fuse;
    sim->csr_valid = false;
    sim->slices_ordered = false;
    sim->layout_valid = false;
    if (defer_host_map && fuse && !(h.host_recombination_rate > 0.0)) return VL_OK;
    TRY(enqueue_host_map(sim));
    if (h.host_recombination_rate > 0.0 || sim->has_rp_rec) TRY(ensure_ordered_slices(sim));
    return VL_OK;
}
static int enqueue_mutate(vl_sim *sim) {
    const DevState &h = sim->h;
    drop_prefill(sim);
    sim->pop_fit_valid = false;  // recombinants and mutants replace haplotypes of the current population
    const bool replay = sim->has_rp_mut || sim->has_rp_rec;
    ReplayRecomb rr{sim->rp_rec_n, sim->rp_rec_host, sim->rp_rec_i, sim->rp_rec_j, sim->rp_rec_s0, sim->rp_rec_s1, sim->rp_rec_which};
    ReplayMut rm{sim->rp_mut_off, sim->rp_mut_sites, sim->rp_mut_bases};
    if (replay) {
        if (sim->has_rp_rec && sim->rp_rec_n > 0) {
            TRY(ensure_ordered_slices(sim));
            LAUNCH(sim, k_recombine, grid_for(sim, h.H), TPB, 0, sim->d, rr, 1);
        }
    } else if (h.host_recombination_rate > 0.0) {
        TRY(ensure_ordered_slices(sim));
        LAUNCH(sim, k_recombine, grid_for(sim, h.H), TPB, 0, sim->d, rr, 0);
    }
    if (replay && !sim->has_rp_mut) {
        // a replayed generation without mutation log means "no mutations"
        sim->has_rp_rec = false;
        return VL_OK;
    }
    if (sim->has_rp_mut || !sim->mutcounts_ready) {
        if (sim->bucket_pending) TRY(enqueue_bucket_host_map(sim));
        if (!sim->csr_valid) return fail(sim, VL_ERR_IMPLEMENTATION, "mutate needs the infections: call vl_infect first");
        TRY(zero_field32(sim, FIELD_OFF(n_mut_list)));
        LAUNCH(sim, k_mutation_counts, grid_for(sim, sim->n_bound), TPB, 0, sim->d, rm, sim->has_rp_mut ? 1 : 0);
    }
    LAUNCH_MUTATE(sim, grid_for(sim, sim->n_bound, 4), TPB, mutate_smem_bytes(h.L), sim->d, rm, sim->has_rp_mut ? 1 : 0, sim->csr_valid ? 1 : (sim->bucket_pending ? 2 : 0), 0, 0);
    sim->pop_fit_valid = false;  // haplotypes of the current population changed
    sim->mutcounts_ready = false;
    sim->has_rp_mut = false;
    sim->has_rp_rec = false;
    return VL_OK;
}
static int enqueue_replicate(vl_sim *sim, int store_lambda = 1, bool tiles_planned = false) {
    const DevState &h = sim->h;
    if (!sim->csr_valid) return fail(sim, VL_ERR_IMPLEMENTATION, "replicate needs the host map: call vl_infect first");
    TRY(zero_replicate_counters(sim));
    const uint64_t *rpo = sim->has_rp_off ? sim->rp_off : nullptr;
    const bool small = h.host_r0 < 12.0;
    // Fused kernel (one pass, whole hosts per CTA) when a host tile is expected to fit a CTA's staging area; the general
    // kernels follow guarded by DevState::fused_failed unless the infection was uniform (then a failure is an error).
    const uint32_t ht = host_tile_size(sim->n_bound, h.H);
    const uint64_t n_host_tiles = ht ? (h.H + ht - 1) / ht : 0;
    const bool fused = !(sim->flags & VL_FLAG_GENERAL_REPLICATE) && ht > 0 && n_host_tiles + 1 <= tile_capacity(h.cap_inf);
    const bool fused_only = fused && sim->uniform_infection;
    int guarded = 0;
    if (fused) {
        if (!tiles_planned) LAUNCH(sim, k_plan_host_tiles, 1, 1, 0, sim->d);  // (k_begin_generation did it for the fused generation)
        const uint32_t per = n_host_tiles >= (uint64_t)sim->n_sms * 12 ? 2u : 1u;  // consecutive tiles per CTA (prefetch of the second)
        const unsigned g = (unsigned)((n_host_tiles + per - 1) / per);
        if (small) LAUNCH(sim, k_replicate_fused<true>, g, REPL_TPB, 0, sim->d, rpo, store_lambda, fused_only ? 1 : 0, per);
        else LAUNCH(sim, k_replicate_fused<false>, g, REPL_TPB, 0, sim->d, rpo, store_lambda, fused_only ? 1 : 0, per);
        guarded = 1;
    }
    if (!fused_only) {
        LAUNCH(sim, k_fitness, grid_for(sim, sim->n_bound), TPB, 0, sim->d, guarded);
        LAUNCH(sim, k_host_sum, grid_for(sim, h.H), TPB, 0, sim->d, 0, guarded);
        LAUNCH(sim, k_host_sum_medium, grid_for(sim, sim->n_bound / 5 + 1, 2), TPB, 0, sim->d, 0, guarded);
        LAUNCH(sim, k_host_sum_heavy, sim->n_sms, TPB, SMEM_SORT_MAX * 16, sim->d, 0, guarded);
        if (small) LAUNCH(sim, k_offspring<true>, grid_tiles(sim, sim->n_bound, RESAMPLE_TILE), TPB, 0, sim->d, rpo, store_lambda, guarded);
        else LAUNCH(sim, k_offspring<false>, grid_tiles(sim, sim->n_bound, RESAMPLE_TILE), TPB, 0, sim->d, rpo, store_lambda, guarded);
    }
    sim->slices_ordered = true;
    sim->has_rp_off = false;
    sim->layout_valid = true;
    sim->layout_identity = false;
    return VL_OK;
}
// upper bound of the number of resample tiles of the current offspring map (either tile kind)
static inline uint64_t tile_bound(const vl_sim *sim) {
    const uint32_t ht = host_tile_size(sim->n_bound, sim->h.H);
    return std::max<uint64_t>((sim->n_bound + RESAMPLE_TILE - 1) / RESAMPLE_TILE, ht ? (sim->h.H + ht - 1) / ht : 0) + 1;
}
static inline unsigned grid_resample(const vl_sim *sim) { return (unsigned)std::min<uint64_t>(tile_bound(sim), (uint64_t)sim->n_sms * 8); }
// Every resample draws from the offspring map in INFECTANT order, tiles of RESAMPLE_TILE consecutive infectants: an
// offspring map the replicate kernels left at CSR positions is moved there first (one scatter + tile totals).  The draws of
// a generation then do not depend on which kernels built its map (host Vec<usize>, host-map kernels, vl_fused.cuh).
static int canonical_layout(vl_sim *sim) {
    if (!sim->layout_valid) return fail(sim, VL_ERR_IMPLEMENTATION, "no offspring map on the device: call vl_replicate_infectants first or pass one");
    if (sim->layout_identity) return VL_OK;
    uint32_t *tmp = sim->h.rank;  // per-infectant scratch, free once the records are scattered
    zero_device(sim, tmp, (sim->n_bound + 4) * 4);
    LAUNCH(sim, k_unsort_offspring32, grid_for(sim, sim->n_bound), TPB, 0, sim->d, tmp);
    LAUNCH(sim, k_identity_from, grid_for(sim, sim->n_bound), TPB, 0, sim->d, (const uint32_t *)tmp);
    LAUNCH(sim, k_tile_sums, grid_tiles(sim, sim->n_bound, RESAMPLE_TILE), TPB, 0, sim->d);
    sim->csr_valid = false;
    sim->slices_ordered = false;
    sim->layout_identity = true;
    return VL_OK;
}
// how the draws of the device-resident generation (k_resample_pos) land
struct PosOpts {
    bool prefill = false;              // also draw the children's hosts and mutation counts of the next generation
    int own_part = -1;                 // the part written straight into the population being assembled
    const uint64_t *own_off = nullptr; // device: its first slot there (nullptr = 0)
    bool predict_ids = false;          // nothing creates records between these draws and the next k_mutate
    std::function<int()> between;      // enqueued between the plan and the draws (the plan's part sizes are final there)
    int tail = 0;                      // ResamplePosArgs::tail: swap (1) and begin of the next generation (2) by the pass's last CTA
};
// plan + draws of n parts of the current offspring map in one pass each (n <= MAX_PARTS)
static int launch_plan_resample(vl_sim *sim, uint32_t n, const double *factor, int use_amount, const uint64_t *amount, int set_next,
                                uint32_t *const *dst, uint64_t *dst64, int mode, bool tiles = false, double *const *dst_fit = nullptr,
                                const PosOpts *po = nullptr) {
    if (n == 0 || n > (uint32_t)MAX_PARTS) return fail(sim, VL_ERR_ARGUMENT, "between 1 and %d parts per pass", MAX_PARTS);
    if (!tiles) TRY(canonical_layout(sim));
    PlanArgs pa;
    ResampleArgs ra;
    memset(&pa, 0, sizeof pa);
    memset(&ra, 0, sizeof ra);
    pa.n_parts = ra.n_parts = n;
    pa.use_amount = use_amount;
    pa.set_next = set_next;
    pa.force_tree = (sim->flags & VL_FLAG_TREE_PLAN) ? 1 : 0;
    for (uint32_t k = 0; k < n; k++) {
        pa.factor[k] = factor ? factor[k] : 0.0;
        pa.amount[k] = amount ? amount[k] : 0;
        pa.salt[k] = ra.salt[k] = ++sim->salt;
        ra.dst[k] = dst ? dst[k] : nullptr;
    }
    const uint64_t nt_bound = tile_bound(sim);
    const uint32_t smem_tiles = (uint32_t)std::min<uint64_t>(nt_bound, sim->plan_smem_tiles);
    // (plan_sync .. plan_acc are zero here: cleared by the last CTA of the previous plan kernel, vl_kernels.cuh plan_leave)
    // a handful of co-resident CTAs: each draws for its slice of tiles, a counter in DevState is the grid barrier
    // (enough CTAs for the tiles, and for the draws a small sample makes one by one: two per thread)
    const uint64_t direct_draws = std::min<uint64_t>(sim->h.max_population, PLAN_FILL_MAX);
    // (several parts: more (part, tile) pairs to draw for, up to the 32 slices DevState::plan_part has room for)
    const uint64_t max_ctas = (n >= 4 || nt_bound > 6000) ? 2 * PLAN_CTAS : PLAN_CTAS;
    const unsigned ctas = (unsigned)std::min<uint64_t>(max_ctas, std::max<uint64_t>(std::max<uint64_t>(1, nt_bound * n / 256), direct_draws / (2 * PLAN_THREADS)));
    LAUNCH(sim, k_plan_resample, ctas, PLAN_THREADS, smem_tiles * 8, sim->d, pa, smem_tiles);
    if (tiles) {  // offspring counts in infectant order next to hap_id[] / pop_fit[] (vl_fused.cuh): haplotype and mapped fitness of the drawn parents
        if (mode != 0) return fail(sim, VL_ERR_IMPLEMENTATION, "the device-resident generation serves haplotype draws only");
        ResamplePosArgs rt;
        memset(&rt, 0, sizeof rt);
        rt.n_parts = n;
        for (uint32_t k = 0; k < n; k++) { rt.salt[k] = ra.salt[k]; rt.dst[k] = ra.dst[k]; rt.dst_fit[k] = dst_fit ? dst_fit[k] : nullptr; }
        rt.own_part = po ? po->own_part : -1;
        rt.own_off = po ? po->own_off : nullptr;
        rt.predict_ids = po && po->predict_ids ? 1 : 0;
        rt.tail = po ? po->tail : 0;
        for (uint32_t k = 0; k < n; k++)
            if ((int)k != rt.own_part && !rt.dst[k]) return fail(sim, VL_ERR_IMPLEMENTATION, "part %u has no destination", k);
        if (po && po->between) TRY(po->between());
        const unsigned rg = (unsigned)std::min<uint64_t>(tile_bound(sim), (uint64_t)sim->n_sms * 8);
        const bool single = n == 1 && rt.own_part == 0 && !rt.own_off;  // vl_run: the one part is the new population
        const bool prefill = po && po->prefill && rt.own_part >= 0;
        const bool small = sim->h.host_r0 < 12.0;  // (the same condition k_offspring_sum<true> is launched under)
        if (prefill && single && small) LAUNCH(sim, (k_resample_pos<true, false, true>), rg, RT_TPB, CL_CAP * sizeof(uint4), sim->d, rt);
        else if (prefill && single) LAUNCH(sim, (k_resample_pos<true, false>), rg, RT_TPB, CL_CAP * sizeof(uint4), sim->d, rt);
        else if (prefill && small) LAUNCH(sim, (k_resample_pos<true, true, true>), rg, RT_TPB, CL_CAP * sizeof(uint4), sim->d, rt);
        else if (prefill) LAUNCH(sim, (k_resample_pos<true, true>), rg, RT_TPB, CL_CAP * sizeof(uint4), sim->d, rt);
        else if (single) LAUNCH(sim, (k_resample_pos<false, false>), rg, RT_TPB, 0, sim->d, rt);
        else LAUNCH(sim, (k_resample_pos<false, true>), rg, RT_TPB, 0, sim->d, rt);
        return VL_OK;
    }
    LAUNCH(sim, k_resample, grid_resample(sim), TPB, 0, sim->d, ra, dst64, mode);
    return VL_OK;
}
// Population::sample as plan + tile-local draws.  dst == nullptr writes the population being assembled.
static int enqueue_resample(vl_sim *sim, double factor, int use_amount, uint64_t amount, uint32_t *dst, uint64_t *dst64, int mode, int set_next) {
    if (!sim->layout_valid) return fail(sim, VL_ERR_IMPLEMENTATION, "no offspring map on the device: call vl_replicate_infectants first or pass one");
    uint32_t *dsts[1] = {dst};
    return launch_plan_resample(sim, 1, &factor, use_amount, &amount, set_next, dsts, dst64, mode);
}
static int enqueue_gc(vl_sim *sim, int force) {
    const DevState &h = sim->h;
    const uint32_t *active = (const uint32_t *)((const char *)sim->d + FIELD_OFF(gc_active));
    const uint64_t *nw_ptr = (const uint64_t *)((const char *)sim->d + FIELD_OFF(gc_nw));
    const uint64_t gw = h.cap_hap / 32 + 8;
    LAUNCH(sim, k_gc_begin, 1, 1, 0, sim->d, force);
    LAUNCH(sim, k_gc_clear, grid_for(sim, gw), TPB, 0, sim->d, 1, 1);
    LAUNCH(sim, k_gc_mark, grid_for(sim, sim->n_bound), TPB, 0, sim->d, (const uint32_t *)nullptr, (uint64_t)0, 1, 1, 0);
    for (vl_pop *p : sim->pops) LAUNCH(sim, k_gc_mark, grid_for(sim, p->cap), TPB, 0, sim->d, (const uint32_t *)p->ids, p->n, 0, 1, 0);
    if (sim->prefilled) LAUNCH(sim, k_gc_mark_pending, grid_for(sim, sim->n_bound), TPB, 0, sim->d);
    LAUNCH(sim, k_gc_words, grid_for(sim, gw), TPB, 0, sim->d, 1);
    TRY((launch_scan<uint32_t, true>(sim, h.gc_wcnt, h.gc_cbase, nw_ptr, 0, gw, sim->gc_totals + 0, 0, active)));
    TRY((launch_scan<uint64_t, true>(sim, h.gc_wlen, h.gc_lbase, nw_ptr, 0, gw, sim->gc_totals + 1, 0, active)));
    LAUNCH(sim, k_gc_copy, grid_for(sim, gw), TPB, 0, sim->d);
    LAUNCH(sim, k_gc_remap, grid_for(sim, sim->n_bound), TPB, 0, sim->d, (uint32_t *)nullptr, (uint64_t)0, 1, (const uint64_t *)sim->gc_totals);
    for (vl_pop *p : sim->pops) LAUNCH(sim, k_gc_remap, grid_for(sim, p->cap), TPB, 0, sim->d, p->ids, p->n, 0, (const uint64_t *)sim->gc_totals);
    if (sim->prefilled) LAUNCH(sim, k_gc_remap_pending, grid_for(sim, sim->n_bound), TPB, 0, sim->d);
    LAUNCH(sim, k_gc_end, 1, 1, 0, sim->d, sim->gc_totals);
    sim->gens_since_gc = 0;
    return VL_OK;
}
static int maybe_gc(vl_sim *sim) {
    if (++sim->gens_since_gc >= sim->gc_period) return enqueue_gc(sim, 0);
    return VL_OK;
}
static void swap_population(vl_sim *sim) {
    LAUNCH(sim, k_swap_population, 1, 1, 0, sim->d);
    std::swap(sim->h.hap_id, sim->h.hap_id_next);
    sim->pop_fit_valid = false;  // (assembled from parts: the per-infectant fitness array was not filled)
}

VL_API int vl_infect(vl_sim *sim) {
    ENTER(sim);
    return enqueue_infect(sim);
}
VL_API int vl_mutate_infectants(vl_sim *sim) {
    ENTER(sim);
    return enqueue_mutate(sim);
}

// An offspring map handed in by the host is aligned with the population order (Vec<usize>, one entry per
// infectant): it is laid out as identity records so that the same plan/draw kernels consume it.
static int upload_offspring(vl_sim *sim, const uint64_t *offspring, uint64_t n) {
    uint64_t cur = 0;
    TRY(population_len(sim, &cur));
    if (n != cur) return fail(sim, VL_ERR_ARGUMENT, "offspring map has %llu entries, population has %llu", (unsigned long long)n, (unsigned long long)cur);
    if (n) CU(sim, cudaMemcpyAsync(sim->stage64, offspring, n * 8, cudaMemcpyHostToDevice, sim->stream));
    LAUNCH(sim, k_identity_records, grid_for(sim, n), TPB, 0, sim->d);
    LAUNCH(sim, k_u64_to_u32, grid_for(sim, n), TPB, 0, sim->d, (const uint64_t *)sim->stage64, sim->h.offspring, n);
    LAUNCH(sim, k_tile_sums, grid_tiles(sim, n, RESAMPLE_TILE), TPB, 0, sim->d);
    sim->csr_valid = false;
    sim->slices_ordered = false;
    sim->layout_valid = true;
    sim->layout_identity = true;
    sim->uniform_infection = false;
    return VL_OK;
}

VL_API int vl_replicate_infectants(vl_sim *sim, uint64_t *offspring_out, uint64_t n) {
    ENTER(sim);
    TRY(enqueue_replicate(sim));
    if (offspring_out) {
        uint64_t cur = 0;
        TRY(population_len(sim, &cur));
        if (n != cur) return fail(sim, VL_ERR_ARGUMENT, "offspring buffer has %llu entries, population has %llu", (unsigned long long)n, (unsigned long long)cur);
        if (n) CU(sim, cudaMemsetAsync(sim->stage64, 0, n * 8, sim->stream));  // uninfected infectants: 0 offspring
        LAUNCH(sim, k_unsort_offspring, grid_for(sim, n), TPB, 0, sim->d, sim->stage64);
        if (n) CU(sim, cudaMemcpyAsync(offspring_out, sim->stage64, n * 8, cudaMemcpyDeviceToHost, sim->stream));
        CU(sim, cudaStreamSynchronize(sim->stream));
        return check_device_error(sim);
    }
    return VL_OK;
}

VL_API int vl_target_size(vl_sim *sim, const uint64_t *offspring, uint64_t n, double *out) {
    ENTER(sim);
    if (offspring) TRY(upload_offspring(sim, offspring, n));
    if (!sim->layout_valid) return fail(sim, VL_ERR_IMPLEMENTATION, "no offspring map on the device");
    LAUNCH(sim, k_target_size, 1, TPB, 0, sim->d);
    TRY(read_field(sim, FIELD_OFF(target_size), out));
    return check_device_error(sim);
}

// Detached populations come and go every generation of a multi-compartment run: their id buffers are pooled per
// handle.  Every user of a buffer is enqueued on the handle's stream, so reuse in stream order needs no synchronisation.
static vl_pop *new_pop(vl_sim *sim, uint64_t cap) {
    vl_pop *p = new vl_pop();
    p->owner = sim; p->ids = nullptr; p->n = 0; p->cap = cap; p->image = nullptr; p->image_bytes = 0; p->pending_slot = -1;
    const uint64_t want = std::max<uint64_t>(cap, 1);
    int best = -1;
    for (size_t k = 0; k < sim->id_pool.size(); k++)
        if (sim->id_pool[k].cap >= want && (best < 0 || sim->id_pool[k].cap < sim->id_pool[best].cap)) best = (int)k;
    if (best >= 0 && sim->id_pool[best].cap <= 4 * want + 4096) {
        p->ids = sim->id_pool[best].p;
        p->cap_alloc = sim->id_pool[best].cap;
        sim->id_pool.erase(sim->id_pool.begin() + best);
    } else {
        if (cudaMalloc((void **)&p->ids, want * 4) != cudaSuccess) { delete p; return nullptr; }
        p->cap_alloc = want;
    }
    sim->pops.push_back(p);
    return p;
}
static void drop_pop(vl_pop *p) {
    vl_sim *sim = p->owner;
    auto it = std::find(sim->pops.begin(), sim->pops.end(), p);
    if (it != sim->pops.end()) sim->pops.erase(it);
    if (p->ids) {
        if (sim->id_pool.size() < 64) sim->id_pool.push_back({p->ids, p->cap_alloc});
        else cudaFree(p->ids);
    }
    if (p->image) cudaFree(p->image);
    delete p;
}

VL_API int vl_sample_indices(vl_sim *sim, const uint64_t *offspring, uint64_t n, uint64_t amount, uint64_t *indices_out) {
    ENTER(sim);
    if (offspring) TRY(upload_offspring(sim, offspring, n));
    uint64_t cur = 0;
    TRY(population_len(sim, &cur));
    if (cur == 0) return VL_OK;  // sample_indices returns an empty vector for an empty map (src/simulation.rs:506-508)
    uint64_t *dst = sim->stage64;
    bool temp = false;
    if (amount > sim->h.cap_inf) { CU(sim, cudaMalloc((void **)&dst, amount * 8)); temp = true; }
    TRY(enqueue_resample(sim, 0.0, 1, amount, nullptr, dst, 1, 0));
    if (amount) CU(sim, cudaMemcpyAsync(indices_out, dst, amount * 8, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    if (temp) cudaFree(dst);
    return check_device_error(sim);
}

// enqueue plan + draws of n detached populations (batches of MAX_PARTS share one pass over the offspring map); each
// capacity bounds its size, the sizes themselves arrive in pinned slots and are read by finalize_pop after a stream
// synchronisation
static int enqueue_subsamples(vl_sim *sim, uint32_t n, const double *factor, int use_amount, const uint64_t *amount, vl_pop **out) {
    if (!sim->layout_valid) return fail(sim, VL_ERR_IMPLEMENTATION, "no offspring map on the device: call vl_replicate_infectants first or pass one");
    for (uint32_t k = 0; k < n; k++) out[k] = nullptr;
    for (uint32_t k0 = 0; k0 < n; k0 += MAX_PARTS) {
        const uint32_t m = std::min<uint32_t>(MAX_PARTS, n - k0);
        uint32_t *dsts[MAX_PARTS];
        for (uint32_t k = 0; k < m; k++) {
            double bound = use_amount ? (double)amount[k0 + k] : floor(factor[k0 + k] * (double)sim->h.max_population) + 1.0;
            if (!(bound >= 0.0)) bound = 0.0;
            if (bound > 4.0e9) return fail(sim, VL_ERR_CAPACITY, "subsample bound too large");
            vl_pop *p = new_pop(sim, (uint64_t)bound + 1);
            if (!p) return fail(sim, VL_ERR_CUDA, "out of device memory for a detached population of %.0f", bound);
            out[k0 + k] = p;
            dsts[k] = p->ids;
        }
        TRY(launch_plan_resample(sim, m, factor ? factor + k0 : nullptr, use_amount, amount ? amount + k0 : nullptr, 0, dsts, nullptr, 0));
        if (sim->next_slot + (int)m >= PIN_SLOTS) sim->next_slot = 128;
        const int slot = sim->next_slot;
        sim->next_slot += (int)m;
        CU(sim, cudaMemcpyAsync(sim->pin + slot, (const char *)sim->d + FIELD_OFF(plan_size), 8 * m, cudaMemcpyDeviceToHost, sim->stream));
        for (uint32_t k = 0; k < m; k++) out[k0 + k]->pending_slot = slot + (int)k;
    }
    return VL_OK;
}
static int enqueue_subsample(vl_sim *sim, double factor, int use_amount, uint64_t amount, vl_pop **out) {
    return enqueue_subsamples(sim, 1, &factor, use_amount, &amount, out);
}
static void finalize_pop(vl_pop *p) {
    if (p->pending_slot >= 0) {
        p->n = std::min<uint64_t>(p->owner->pin[p->pending_slot], p->cap);
        p->pending_slot = -1;
    }
}

VL_API int vl_subsample_population(vl_sim *sim, const uint64_t *offspring, uint64_t n, double factor, vl_pop **out) {
    ENTER(sim);
    if (!out) return fail(sim, VL_ERR_ARGUMENT, "null out");
    if (offspring) TRY(upload_offspring(sim, offspring, n));
    TRY(enqueue_subsample(sim, factor, 0, 0, out));
    CU(sim, cudaStreamSynchronize(sim->stream));
    finalize_pop(*out);
    int r = check_device_error(sim);
    if (r != VL_OK) { drop_pop(*out); *out = nullptr; }
    return r;
}

VL_API int vl_select(vl_sim *sim, const uint64_t *indices, uint64_t n, vl_pop **out) {
    ENTER(sim);
    if (!out) return fail(sim, VL_ERR_ARGUMENT, "null out");
    vl_pop *p = new_pop(sim, n);
    if (!p) return fail(sim, VL_ERR_CUDA, "out of device memory");
    uint64_t *src = sim->stage64;
    bool temp = false;
    if (n > sim->h.cap_inf) { CU(sim, cudaMalloc((void **)&src, n * 8)); temp = true; }
    if (n) CU(sim, cudaMemcpyAsync(src, indices, n * 8, cudaMemcpyHostToDevice, sim->stream));
    LAUNCH(sim, k_select, grid_for(sim, n), TPB, 0, sim->d, src, n, p->ids);
    p->n = n;
    CU(sim, cudaStreamSynchronize(sim->stream));
    if (temp) cudaFree(src);
    int r = check_device_error(sim);
    if (r != VL_OK) { drop_pop(p); return r; }
    *out = p;
    return VL_OK;
}

VL_API int vl_pop_len(const vl_pop *pop, uint64_t *out) {
    if (!pop || !out) return fail(nullptr, VL_ERR_ARGUMENT, "null argument");
    *out = pop->n;
    return VL_OK;
}
VL_API int vl_pop_free(vl_pop *pop) {
    if (!pop) return VL_OK;
    vl_sim *sim = pop->owner;
    ENTER(sim);
    if (pop->image) CU(sim, cudaStreamSynchronize(sim->stream));  // the image is freed, the id buffer goes back to the pool
    drop_pop(pop);
    return VL_OK;
}

// ---------------------------------------------------------------------------------------------- migrants
static int pack_pop(vl_sim *sim, vl_pop *pop) {
    if (pop->image) return VL_OK;
    const DevState &h = sim->h;
    const uint64_t gw = h.cap_hap / 32 + 8;
    const uint64_t *nw_ptr = (const uint64_t *)((const char *)sim->d + FIELD_OFF(gc_nw));
    LAUNCH(sim, k_gc_clear, grid_for(sim, gw), TPB, 0, sim->d, 0, 0);
    LAUNCH(sim, k_gc_mark, grid_for(sim, pop->n), TPB, 0, sim->d, (const uint32_t *)pop->ids, pop->n, 0, 0, 1);
    LAUNCH(sim, k_gc_words, grid_for(sim, gw), TPB, 0, sim->d, 0);
    TRY((launch_scan<uint32_t, true>(sim, h.gc_wcnt, h.gc_cbase, nw_ptr, 0, gw, sim->gc_totals + 2, 0, nullptr)));
    TRY((launch_scan<uint64_t, true>(sim, h.gc_wlen, h.gc_lbase, nw_ptr, 0, gw, sim->gc_totals + 3, 0, nullptr)));
    CU(sim, cudaMemcpyAsync(sim->pin, sim->gc_totals + 2, 16, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    const uint64_t n_unique = sim->pin[0], n_words = sim->pin[1];
    PackHeader hd = make_pack_header(pop->n, n_unique, n_words, h.n_providers);
    CU(sim, cudaMalloc((void **)&pop->image, hd.bytes));
    pop->image_bytes = hd.bytes;
    LAUNCH(sim, k_pack_copy, grid_for(sim, std::max<uint64_t>(gw, pop->n)), TPB, 0, sim->d, (const uint32_t *)pop->ids, pop->image, hd);
    return VL_OK;
}
VL_API int vl_pop_pack_device(vl_sim *origin, vl_pop *pop, void **image_device, uint64_t *bytes) {
    ENTER(origin);
    if (!pop || pop->owner != origin || !image_device || !bytes) return fail(origin, VL_ERR_ARGUMENT, "population does not belong to this handle");
    TRY(pack_pop(origin, pop));
    CU(origin, cudaStreamSynchronize(origin->stream));
    *image_device = pop->image;
    *bytes = pop->image_bytes;
    return check_device_error(origin);
}
// unpack into dst (a detached population's ids) or, when dst == nullptr, into the population being assembled
static int unpack_image(vl_sim *sim, const uint8_t *image, uint64_t bytes, uint32_t *dst, uint64_t dst_off, uint64_t *n_pop_out) {
    PackHeader hd;
    if (bytes < sizeof(PackHeader)) return fail(sim, VL_ERR_ARGUMENT, "packed image too small");
    CU(sim, cudaMemcpy(&hd, image, sizeof(PackHeader), cudaMemcpyDeviceToHost));
    if (hd.magic != PACK_MAGIC || hd.bytes > bytes) return fail(sim, VL_ERR_ARGUMENT, "not a packed population image");
    if (hd.n_providers != sim->h.n_providers) return fail(sim, VL_ERR_ARGUMENT, "packed image has %llu fitness providers, handle has %u", (unsigned long long)hd.n_providers, sim->h.n_providers);
    LAUNCH(sim, k_reserve_bulk, 1, 1, 0, sim->d, hd.n_unique, hd.n_words, sim->gc_totals + 4);
    LAUNCH(sim, k_unpack, grid_for(sim, std::max<uint64_t>(std::max<uint64_t>(hd.n_unique, hd.n_words), hd.n_pop)), TPB, 0, sim->d, image, hd,
           (const uint64_t *)(sim->gc_totals + 4), dst, dst_off);
    *n_pop_out = hd.n_pop;
    return VL_OK;
}
VL_API int vl_pop_unpack_device(vl_sim *target, const void *image_device, uint64_t bytes, vl_pop **out) {
    ENTER(target);
    if (!image_device || !out) return fail(target, VL_ERR_ARGUMENT, "null argument");
    PackHeader hd;
    if (bytes < sizeof(PackHeader)) return fail(target, VL_ERR_ARGUMENT, "packed image too small");
    CU(target, cudaMemcpy(&hd, image_device, sizeof(PackHeader), cudaMemcpyDeviceToHost));
    if (hd.magic != PACK_MAGIC) return fail(target, VL_ERR_ARGUMENT, "not a packed population image");
    vl_pop *p = new_pop(target, hd.n_pop);
    if (!p) return fail(target, VL_ERR_CUDA, "out of device memory");
    uint64_t n_pop = 0;
    int r = unpack_image(target, (const uint8_t *)image_device, bytes, p->ids, 0, &n_pop);
    if (r == VL_OK) { CU(target, cudaStreamSynchronize(target->stream)); r = check_device_error(target); }
    if (r != VL_OK) { drop_pop(p); return r; }
    p->n = n_pop;
    *out = p;
    return VL_OK;
}

// ---- migrants between processes: flat per-migrant images (vl_migrate.cuh), moved by the caller's collective
static int ensure_buf(vl_sim *sim, vl_sim::DevBuf &b, uint64_t bytes) {
    if (bytes <= b.bytes) return VL_OK;
    CU(sim, cudaStreamSynchronize(sim->stream));
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = bytes + bytes / 2 + 256;
    CU(sim, cudaMalloc(&b.p, b.bytes));
    return VL_OK;
}
VL_API int vl_migrants_export_device(vl_sim *sim, vl_pop *const *parts, uint32_t n_parts, vl_migrant_image *image,
                                     uint64_t *migrants_per_part, uint64_t *words_per_part) {
    ENTER(sim);
    if (!image || (n_parts && (!parts || !migrants_per_part || !words_per_part))) return fail(sim, VL_ERR_ARGUMENT, "null argument");
    if (n_parts > 60) return fail(sim, VL_ERR_ARGUMENT, "too many parts");
    bool pending = false;
    for (uint32_t k = 0; k < n_parts; k++) pending = pending || (parts[k] && parts[k]->pending_slot >= 0);
    if (pending) CU(sim, cudaStreamSynchronize(sim->stream));
    uint64_t M = 0;
    for (uint32_t k = 0; k < n_parts; k++) {
        if (!parts[k] || parts[k]->owner != sim) return fail(sim, VL_ERR_ARGUMENT, "part %u does not belong to this handle", k);
        finalize_pop(parts[k]);
        migrants_per_part[k] = parts[k]->n;
        M += parts[k]->n;
    }
    const uint64_t P = sim->h.n_providers;
    TRY(ensure_buf(sim, sim->mig[0], (M + 1) * 4));
    TRY(ensure_buf(sim, sim->mig[1], (M + 1) * 4));
    TRY(ensure_buf(sim, sim->mig[2], (M * P + 1) * 8));
    TRY(ensure_buf(sim, sim->mig[3], (M + 2) * 8));
    uint32_t *len = (uint32_t *)sim->mig[0].p, *words = (uint32_t *)sim->mig[1].p;
    double *raw = (double *)sim->mig[2].p;
    uint64_t *off = (uint64_t *)sim->mig[3].p;
    uint64_t m0 = 0;
    for (uint32_t k = 0; k < n_parts; k++) {
        if (parts[k]->n) LAUNCH(sim, k_mig_lens, grid_for(sim, parts[k]->n), TPB, 0, sim->d, (const uint32_t *)parts[k]->ids, parts[k]->n, m0, len, words, raw);
        m0 += parts[k]->n;
    }
    TRY((launch_scan<uint64_t, true>(sim, words, off, nullptr, M, M, sim->gc_totals + 10, 1, nullptr)));
    m0 = 0;
    for (uint32_t k = 0; k < n_parts; k++) {  // word offset of every part boundary
        CU(sim, cudaMemcpyAsync(sim->pin + 64 + k, off + m0, 8, cudaMemcpyDeviceToHost, sim->stream));
        m0 += parts[k]->n;
    }
    CU(sim, cudaMemcpyAsync(sim->pin + 64 + n_parts, sim->gc_totals + 10, 8, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    const uint64_t W = sim->pin[64 + n_parts];
    for (uint32_t k = 0; k < n_parts; k++) words_per_part[k] = sim->pin[64 + k + 1] - sim->pin[64 + k];
    TRY(ensure_buf(sim, sim->mig[4], (W + 1) * 4));
    uint32_t *entries = (uint32_t *)sim->mig[4].p;
    m0 = 0;
    for (uint32_t k = 0; k < n_parts; k++) {
        if (parts[k]->n) LAUNCH(sim, k_mig_copy, grid_for(sim, parts[k]->n), TPB, 0, sim->d, (const uint32_t *)parts[k]->ids, parts[k]->n, m0, (const uint64_t *)off, entries);
        m0 += parts[k]->n;
    }
    CU(sim, cudaStreamSynchronize(sim->stream));
    image->len_device = len;
    image->raw_device = raw;
    image->entries_device = entries;
    image->n_migrants = M;
    image->n_words = W;
    image->n_providers = P;
    image->cap_migrants = std::min<uint64_t>(sim->mig[0].bytes / 4, P ? sim->mig[2].bytes / (8 * P) : ~0ull);
    image->cap_words = sim->mig[4].bytes / 4;
    return check_device_error(sim);
}
VL_API int vl_migrants_import_device(vl_sim *sim, const uint32_t *len_device, const double *raw_device, const uint32_t *entries_device,
                                     uint64_t M, uint64_t W, vl_pop **out) {
    ENTER(sim);
    if (!out || (M && (!len_device || (sim->h.n_providers && !raw_device))) || (W && !entries_device)) return fail(sim, VL_ERR_ARGUMENT, "null argument");
    vl_pop *p = new_pop(sim, M);
    if (!p) return fail(sim, VL_ERR_CUDA, "out of device memory");
    int r = VL_OK;
    if (M) {
        r = ensure_buf(sim, sim->mig[5], (M + 1) * 4);
        if (r == VL_OK) r = ensure_buf(sim, sim->mig[6], (M + 1) * 4);
        if (r == VL_OK) r = ensure_buf(sim, sim->mig[7], (M + 2) * 4);
        if (r == VL_OK) r = ensure_buf(sim, sim->mig[8], (M + 2) * 8);
        if (r != VL_OK) { drop_pop(p); return r; }
        uint32_t *words = (uint32_t *)sim->mig[5].p, *rec = (uint32_t *)sim->mig[6].p, *rec_off = (uint32_t *)sim->mig[7].p;
        uint64_t *word_off = (uint64_t *)sim->mig[8].p;
        LAUNCH(sim, k_mig_words, grid_for(sim, M), TPB, 0, len_device, M, words, rec);
        r = launch_scan<uint32_t, true>(sim, rec, rec_off, nullptr, M, M, sim->gc_totals + 12, 0, nullptr);
        if (r == VL_OK) r = launch_scan<uint64_t, true>(sim, words, word_off, nullptr, M, M, sim->gc_totals + 13, 0, nullptr);
        if (r != VL_OK) { drop_pop(p); return r; }
        LAUNCH(sim, k_mig_reserve, 1, 1, 0, sim->d, (const uint64_t *)(sim->gc_totals + 12), W, sim->gc_totals + 14);
        LAUNCH(sim, k_mig_import, grid_for(sim, M), TPB, 0, sim->d, len_device, raw_device, entries_device, M, (const uint32_t *)rec_off,
               (const uint64_t *)word_off, (const uint64_t *)(sim->gc_totals + 14), p->ids);
    }
    p->n = M;
    cudaError_t ce = cudaStreamSynchronize(sim->stream);
    if (ce != cudaSuccess) { drop_pop(p); return fail(sim, VL_ERR_CUDA, "import failed: %s", cudaGetErrorString(ce)); }
    r = check_device_error(sim);
    if (r != VL_OK) { drop_pop(p); return r; }
    *out = p;
    return VL_OK;
}
// several subsamples of the same offspring map with one synchronisation (the per-target parts of one origin,
// src/runner.rs:404-416)
VL_API int vl_subsample_populations(vl_sim *sim, const double *factors, uint32_t n, vl_pop **out) {
    ENTER(sim);
    if (!factors || !out) return fail(sim, VL_ERR_ARGUMENT, "null argument");
    if (n >= (uint32_t)PIN_SLOTS - 16) return fail(sim, VL_ERR_ARGUMENT, "too many parts");
    int r = enqueue_subsamples(sim, n, factors, 0, nullptr, out);
    const uint32_t made = n;
    cudaError_t ce = cudaStreamSynchronize(sim->stream);
    if (r == VL_OK && ce != cudaSuccess) r = fail(sim, VL_ERR_CUDA, "subsample failed: %s", cudaGetErrorString(ce));
    if (r == VL_OK) r = check_device_error(sim);
    for (uint32_t k = 0; k < made; k++) {
        if (!out[k]) continue;
        if (r == VL_OK) finalize_pop(out[k]);
        else { drop_pop(out[k]); out[k] = nullptr; }
    }
    return r;
}

VL_API int vl_set_population(vl_sim *sim, vl_pop *const *parts, uint32_t n_parts) {
    ENTER(sim);
    drop_prefill(sim);
    uint64_t total = 0;
    for (uint32_t k = 0; k < n_parts; k++) {
        if (!parts[k]) return fail(sim, VL_ERR_ARGUMENT, "null part");
        finalize_pop(parts[k]);
        total += parts[k]->n;
    }
    if (total > sim->h.cap_inf) return fail(sim, VL_ERR_CAPACITY, "assembled population of %llu exceeds capacity_infectants %llu", (unsigned long long)total, (unsigned long long)sim->h.cap_inf);
    uint64_t off = 0;
    for (uint32_t k = 0; k < n_parts; k++) {
        vl_pop *p = parts[k];
        if (p->n == 0) continue;
        if (p->owner == sim) {
            LAUNCH(sim, k_copy_ids, grid_for(sim, p->n), TPB, 0, sim->d, (const uint32_t *)p->ids, p->n, off);
        } else {
            vl_sim *o = p->owner;
            {
                std::lock_guard<std::recursive_mutex> lk(o->mu);
                CU(sim, cudaSetDevice(o->device));
                int r = pack_pop(o, p);
                if (r == VL_OK) { cudaStreamSynchronize(o->stream); r = check_device_error(o); }
                if (r != VL_OK) { cudaSetDevice(sim->device); return fail(sim, r, "packing migrants failed: %s", o->err); }
                CU(sim, cudaSetDevice(sim->device));
            }
            const uint8_t *image = p->image;
            uint8_t *local = nullptr;
            if (o->device != sim->device) {
                CU(sim, cudaMalloc((void **)&local, p->image_bytes));
                CU(sim, cudaMemcpyPeer(local, sim->device, p->image, o->device, p->image_bytes));
                image = local;
            }
            uint64_t n_pop = 0;
            int r = unpack_image(sim, image, p->image_bytes, nullptr, off, &n_pop);
            CU(sim, cudaStreamSynchronize(sim->stream));
            if (local) cudaFree(local);
            TRY(r);
        }
        off += p->n;
    }
    LAUNCH(sim, k_set_next_n, 1, 1, 0, sim->d, total);
    swap_population(sim);
    CU(sim, cudaStreamSynchronize(sim->stream));
    for (uint32_t k = 0; k < n_parts; k++) {  // consumes the parts (Population is moved in, src/simulation.rs:340-345)
        vl_sim *o = parts[k]->owner;
        std::lock_guard<std::recursive_mutex> lk(o->mu);
        drop_pop(parts[k]);
    }
    sim->n_known = total;
    sim->n_bound = total;
    sim->n_valid = true;
    sim->csr_valid = false;
    sim->slices_ordered = false;
    sim->layout_valid = false;
    sim->mutcounts_ready = false;
    TRY(check_device_error(sim));
    return maybe_gc(sim);
}

// ---------------------------------------------------------------------------------------------- fused loops
// increment_generation + infect + mutate_infectants + replicate_infectants with the offspring map left on the device
// (src/runner.rs:382-399; the first half of Simulation::next_generation, src/simulation.rs:200-214)
static int enqueue_generation_front(vl_sim *sim) {
    const DevState &h = sim->h;
    drop_prefill(sim);
    const bool replay = sim->has_rp_inf || sim->has_rp_mut || sim->has_rp_rec || sim->has_rp_off;
    sim->salt = 0;
    const bool fast_flow = !(sim->experiment & 1) && !replay && h.binv.ok && !(h.host_recombination_rate > 0.0) && buckets_possible(sim) &&
                           h.cap_tmp >= (uint64_t)h.n_buckets * bucket_stride(sim->n_bound, h.H);
    if (fast_flow) {
        // uniform infection (every infectant is infected, hosts iid uniform): the mutation count of an infectant does not
        // depend on its host, so mutate first and build the host map from the already-mutated population
        LAUNCH(sim, k_begin_generation, 1, 1024, 0, sim->d);
        LAUNCH(sim, k_mutcount_fast, grid_tiles(sim, sim->n_bound, (uint64_t)TPB * MC_ITEMS, 8), TPB, 0, sim->d);
        LAUNCH_MUTATE(sim, grid_for(sim, sim->n_bound / 4 + 1, 8), TPB, mutate_smem_bytes(h.L), sim->d, ReplayMut{nullptr, nullptr, nullptr}, 0, 0, 1, 0);
        const int btpb = (sim->experiment & 2) ? 512 : 1024;
        const unsigned chunks = (unsigned)std::max<uint64_t>(1, (sim->n_bound + btpb * BUCKET_ITEMS - 1) / (btpb * BUCKET_ITEMS));
        if (btpb == 1024) LAUNCH(sim, (k_infect_bucket<false, false, 1024>), chunks, 1024, 0, sim->d);
        else LAUNCH(sim, (k_infect_bucket<false, false, 512>), chunks, 512, 0, sim->d);
        sim->uniform_infection = true;
        sim->mutcounts_ready = false;
        TRY(enqueue_bucket_host_map(sim));
    } else {
        LAUNCH(sim, k_tick, 1, 1, 0, sim->d);
        TRY(enqueue_infect(sim, true));
        TRY(enqueue_mutate(sim));
        if (!sim->csr_valid) TRY(sim->bucket_pending ? enqueue_bucket_host_map(sim) : enqueue_host_map(sim));
    }
    TRY(enqueue_replicate(sim, 0, fast_flow));
    return VL_OK;
}
// ---- the device-resident generation without a host map (vl_fused.cuh)
static bool fused_now(const vl_sim *sim) {
    const bool replay = sim->has_rp_inf || sim->has_rp_mut || sim->has_rp_rec || sim->has_rp_off;
    return !replay && fused_possible(sim);
}
__global__ void __launch_bounds__(TPB) k_zero_hsum_current(DevState *st) {
    double *p = st->hsum[st->hsum_parity];
    const uint64_t H = st->H;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < H; k += (uint64_t)gridDim.x * blockDim.x) p[k] = 0.0;
}
// prefill_next: this generation's resample also draws the next generation's hosts and mutation counts
static int enqueue_generation_fused(vl_sim *sim, bool prefill_next) {
    const DevState &h = sim->h;
    sim->salt = 0;
    const bool prefilled = sim->prefilled;
    if (sim->begun_ahead && !prefilled) return fail(sim, VL_ERR_IMPLEMENTATION, "a generation was begun ahead without its prepared infection");
    if (!sim->begun_ahead) LAUNCH(sim, k_begin_generation_fused, 1, 1, 0, sim->d, prefilled ? 1 : 0);  // (else: the previous resample's last CTA did it)
    if (!prefilled) {
        if (sim->hsum_dirty) LAUNCH(sim, k_zero_hsum_current, grid_for(sim, h.H), TPB, 0, sim->d);
        sim->hsum_dirty = false;
        const unsigned ig = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (sim->n_bound + RESAMPLE_TILE - 1) / RESAMPLE_TILE), (uint64_t)sim->n_sms * 32);
        if (sim->pop_fit_valid) LAUNCH(sim, k_infect_sum<false>, ig, IS_TPB, 0, sim->d);
        else LAUNCH(sim, k_infect_sum<true>, ig, IS_TPB, 0, sim->d);
    }
    LAUNCH_MUTATE_FUSED(sim, grid_for(sim, sim->n_bound / 4 + 1, 8), TPB, mutate_smem_bytes(h.L), sim->d, ReplayMut{nullptr, nullptr, nullptr}, 0,
           prefilled && sim->prefill_patch ? 4 : 3, 1, prefilled && sim->prefill_own_done ? 2 : 0);
    const unsigned og = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (sim->n_bound + RESAMPLE_TILE - 1) / RESAMPLE_TILE), (uint64_t)sim->n_sms * 32);
    if (h.host_r0 < 12.0) LAUNCH(sim, k_offspring_sum<true>, og, OS_TPB, 0, sim->d);
    else LAUNCH(sim, k_offspring_sum<false>, og, OS_TPB, 0, sim->d);
    const double one = 1.0;
    uint32_t *dsts[1] = {nullptr};
    PosOpts po;
    po.prefill = prefill_next;
    po.own_part = 0;
    po.predict_ids = true;
    po.tail = 1 | (prefill_next ? 2 : 0);  // the resample's last CTA installs the population and, if it prepared it, begins the next generation
    TRY(launch_plan_resample(sim, 1, &one, 0, nullptr, 1, dsts, nullptr, 0, true, nullptr, &po));
    sim->begun_ahead = prefill_next;
    sim->prefilled = prefill_next;
    sim->prefill_patch = false;
    sim->prefill_own_done = false;
    sim->uniform_infection = true;
    sim->mutcounts_ready = false;
    sim->bucket_pending = false;
    return VL_OK;
}
// increment_generation + infect + mutate_infectants + replicate_infectants on the kernels of vl_fused.cuh, offspring map
// left on the device in infectant order (what vl_exchange_begin enqueues when the compartment qualifies)
static int enqueue_generation_front_fused(vl_sim *sim) {
    const DevState &h = sim->h;
    sim->salt = 0;
    const bool prefilled = sim->prefilled;  // (the previous exchange generation of the same vl_exchange_run call prepared this one)
    LAUNCH(sim, k_begin_generation_fused, 1, 1, 0, sim->d, prefilled ? 1 : 0);
    const unsigned ig = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (sim->n_bound + RESAMPLE_TILE - 1) / RESAMPLE_TILE), (uint64_t)sim->n_sms * 32);
    if (!prefilled) {
        if (sim->hsum_dirty) LAUNCH(sim, k_zero_hsum_current, grid_for(sim, h.H), TPB, 0, sim->d);
        sim->hsum_dirty = false;
        if (sim->pop_fit_valid) LAUNCH(sim, k_infect_sum<false>, ig, IS_TPB, 0, sim->d);
        else LAUNCH(sim, k_infect_sum<true>, ig, IS_TPB, 0, sim->d);
    }
    LAUNCH_MUTATE_FUSED(sim, grid_for(sim, sim->n_bound / 4 + 1, 8), TPB, mutate_smem_bytes(h.L), sim->d, ReplayMut{nullptr, nullptr, nullptr}, 0,
           prefilled && sim->prefill_patch ? 4 : 3, 1, prefilled && sim->prefill_own_done ? 2 : 0);
    sim->prefilled = false;
    if (h.host_r0 < 12.0) LAUNCH(sim, k_offspring_sum<true>, ig, OS_TPB, 0, sim->d);
    else LAUNCH(sim, k_offspring_sum<false>, ig, OS_TPB, 0, sim->d);
    sim->uniform_infection = true;
    sim->mutcounts_ready = false;
    sim->bucket_pending = false;
    sim->csr_valid = false;
    sim->layout_valid = false;
    return VL_OK;
}
// the kernels of one whole generation (everything but the occasional compaction)
static int enqueue_generation_kernels(vl_sim *sim, bool tiles, bool prefill_next) {
    if (tiles) return enqueue_generation_fused(sim, prefill_next);
    drop_prefill(sim);
    TRY(enqueue_generation_front(sim));
    TRY(enqueue_resample(sim, 1.0, 0, 0, nullptr, nullptr, 0, 1));
    LAUNCH(sim, k_swap_population, 1, 1, 0, sim->d);
    return VL_OK;
}
constexpr uint64_t GRAPH_MAX_POPULATION = 4u << 20;  // above this a generation is milliseconds of kernels, launches do not matter
static bool graph_wanted(const vl_sim *sim) {
    const DevState &h = sim->h;
    const bool replay = sim->has_rp_inf || sim->has_rp_mut || sim->has_rp_rec || sim->has_rp_off;
    return !sim->graph_disabled && !sim->profiling && !(sim->experiment & 64) && !replay && (sim->n_bound <= GRAPH_MAX_POPULATION || (sim->experiment & 128)) &&
           sim->n_bound == std::min<uint64_t>(h.cap_inf, h.max_population);  // the bound every later generation has too
}
// more_follow: the caller enqueues another generation right after this one (vl_run): this generation's resample may then
// draw the next generation's hosts and mutation counts.  Nothing of that prepared state outlives the call.
static int enqueue_generation(vl_sim *sim, bool more_follow) {
    // Simulation::next_generation (src/simulation.rs:200-218).  An empty population makes every kernel a no-op.
    const bool tiles = fused_now(sim);
    if (!tiles) drop_prefill(sim);
    const bool prefill_next = tiles && more_follow && !(sim->experiment & 32);
    if (graph_wanted(sim)) {
        const DevState &h = sim->h;
        const uint64_t key[5] = {sim->n_bound, h.H, h.cap_inf, (uint64_t)(uintptr_t)h.tmp_rec ^ (uint64_t)(uintptr_t)h.offsets ^ (uint64_t)(uintptr_t)h.hsum[0],
                                 (uint64_t)sim->flags | (tiles ? 1ull << 40 : 0) | (sim->pop_fit_valid ? 1ull << 41 : 0) | (sim->prefilled ? 1ull << 42 : 0) |
                                     (prefill_next ? 1ull << 43 : 0) | (sim->hsum_dirty ? 1ull << 44 : 0) | (sim->begun_ahead ? 1ull << 45 : 0)};
        if (!sim->gen_graph || memcmp(key, sim->graph_key, sizeof key) != 0) {
            if (sim->gen_graph) { cudaGraphExecDestroy(sim->gen_graph); sim->gen_graph = nullptr; }
            cudaGraph_t graph = nullptr;
            const uint64_t l0 = sim->launches;
            std::lock_guard<std::mutex> context_lock(g_context_mu);
            if (cudaStreamBeginCapture(sim->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int rc = enqueue_generation_kernels(sim, tiles, prefill_next);
                const cudaError_t ce = cudaStreamEndCapture(sim->stream, &graph);
                if (rc == VL_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&sim->gen_graph, graph, 0) == cudaSuccess) {
                    memcpy(sim->graph_key, key, sizeof key);
                    sim->graph_launches = sim->launches - l0;
                } else {
                    sim->gen_graph = nullptr;
                    sim->graph_disabled = true;  // whatever made the capture fail would fail again: enqueue directly from now on
                }
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
            } else {
                sim->graph_disabled = true;
                cudaGetLastError();
            }
            sim->launches = l0;  // (a failed capture enqueued nothing: the direct path below runs the generation)
        }
        if (sim->gen_graph) {
            CU(sim, cudaGraphLaunch(sim->gen_graph, sim->stream));
            sim->launches += sim->graph_launches;
            if (tiles) { sim->prefilled = prefill_next; sim->begun_ahead = prefill_next; sim->hsum_dirty = false; }
            sim->salt = 1;
            sim->mutcounts_ready = false;
            sim->bucket_pending = false;
            sim->has_rp_off = false;
        } else {
            TRY(enqueue_generation_kernels(sim, tiles, prefill_next));
        }
    } else {
        TRY(enqueue_generation_kernels(sim, tiles, prefill_next));
    }
    std::swap(sim->h.hap_id, sim->h.hap_id_next);
    if (tiles) std::swap(sim->h.pop_fit, sim->h.pop_fit_next);
    sim->pop_fit_valid = tiles;  // the children inherited their parents' mapped fitness (k_resample_pos)
    sim->n_valid = false;
    sim->n_bound = std::min<uint64_t>(sim->h.cap_inf, sim->h.max_population);
    sim->csr_valid = false;
    sim->slices_ordered = false;
    sim->layout_valid = false;
    return maybe_gc(sim);
}
VL_API int vl_advance_to_offspring(vl_sim *sim) {
    ENTER(sim);
    TRY(enqueue_generation_front(sim));
    CU(sim, cudaGetLastError());
    return VL_OK;
}
VL_API int vl_next_generation(vl_sim *sim) {
    ENTER(sim);
    TRY(enqueue_generation(sim, false));
    CU(sim, cudaGetLastError());
    return VL_OK;
}
VL_API int vl_run(vl_sim *sim, uint64_t n_generations) {
    ENTER(sim);
    for (uint64_t g = 0; g < n_generations; g++) TRY(enqueue_generation(sim, g + 1 < n_generations));
    CU(sim, cudaGetLastError());
    return VL_OK;
}

// ---------------------------------------------------------------------------------------------- peer exchange
// One compartment per process/GPU; migrants are written by the origin straight into the target's receive buffer
// (vl_exchange.cuh).  Create -> exchange the 64-byte IPC handles with whatever the caller has (torch.distributed in
// virolution_b200/distributed.py) -> connect -> vl_exchange_step per generation.
struct vl_exchange {
    vl_sim *sim = nullptr;
    uint32_t rank = 0, world = 0;
    uint64_t wpm = 0, seq = 0;
    std::vector<double> T;         // the matrix of the generation being enqueued (vl_exchange_set_transfer)
    std::vector<double> Tcap;      // the matrix the buffers were sized for (vl_exchange_create): an upper bound of every T, entry by entry
    uint8_t *recv = nullptr;       // my receive buffer (IPC-exported)
    uint64_t recv_bytes = 0;
    void *peer[EX_MAX_WORLD] = {nullptr};
    bool peer_ipc[EX_MAX_WORLD] = {false};
    bool connected = false, pushed = false, begun = false;
    bool sharded = false;         // one compartment split over the ranks: amounts come from the device-side multinomial
    uint64_t max_pop = 0;
    ExArgs args;
    ExScratch *scratch = nullptr;
    uint32_t *part_ids[EX_MAX_WORLD] = {nullptr};
    double *own_fit = nullptr;   // mapped fitness of my own part (device-resident generation)
    bool fused_step = false;     // the generation in flight runs on the kernels of vl_fused.cuh
    bool fast_step = false;      // ... and on the one-pass export / import kernels (own part drawn in place)
    bool prefill_step = false;   // ... and prepares the next generation's infection (vl_exchange_run, all but the last generation)
    bool more_follow = false;
    bool overlap_step = false;   // ... and mutates its own part of the next generation while the migrants travel (side stream)
    cudaStream_t side = nullptr; // export, wait and import of an overlapped generation
    cudaEvent_t ev_drawn = nullptr, ev_imported = nullptr;
    uint32_t *words_out = nullptr, *words_in = nullptr, *recs_in = nullptr, *rec_off = nullptr;
    uint64_t *woff = nullptr, *word_off = nullptr;
    uint32_t *stage = nullptr;  // leaving entries in concatenation order, before they are streamed to the targets
    uint64_t m_out_bound = 0, m_in_bound = 0;
};
struct ExLayout {
    uint64_t tflag_off[2], tot_off[2];  // sharded compartment: world flags + world offspring totals per parity
    uint64_t flag_off[2][EX_MAX_WORLD], header_off[2][EX_MAX_WORLD], len_off[2][EX_MAX_WORLD], raw_off[2][EX_MAX_WORLD], ent_off[2][EX_MAX_WORLD];
    uint64_t eflag_off[2][EX_MAX_WORLD], meta_off[2][EX_MAX_WORLD];
    uint64_t m_cap[EX_MAX_WORLD], w_cap[EX_MAX_WORLD];
    uint64_t bytes;
};
// layout of rank `target`'s receive buffer; every rank computes the same layout for every rank
// migrants one (target <- origin) cell of a sharded compartment can hold: mean max_pop / world^2 with a quarter of slack
// for unequal shard totals plus 8 standard deviations
static uint64_t shard_cell_capacity(uint64_t max_pop, uint32_t world) {
    const double mean = (double)max_pop / ((double)world * (double)world);
    return (uint64_t)(1.25 * mean + 8.0 * sqrt(mean) + 1024.0);
}
static ExLayout ex_layout(uint32_t world, uint32_t target, const double *T, uint64_t max_pop, uint32_t P, uint64_t wpm, bool sharded) {
    ExLayout L;
    memset(&L, 0, sizeof L);
    uint64_t off = 0;
    for (int par = 0; par < 2; par++)
        for (uint32_t o = 0; o < world; o++) { L.flag_off[par][o] = off; off += 8; }
    for (int par = 0; par < 2; par++)
        for (uint32_t o = 0; o < world; o++) { L.eflag_off[par][o] = off; off += 8; }
    for (int par = 0; par < 2; par++) { L.tflag_off[par] = off; off += 8ull * world; }
    for (int par = 0; par < 2; par++) { L.tot_off[par] = off; off += 8ull * world; }
    off = (off + 255) & ~255ull;
    for (uint32_t o = 0; o < world; o++) {
        double m = sharded ? (double)shard_cell_capacity(max_pop, world) : floor(T[(size_t)target * world + o] * (double)max_pop) + 1.0;
        if (!(m >= 1.0)) m = 1.0;
        L.m_cap[o] = o == target ? 0 : (uint64_t)m;
        L.w_cap[o] = L.m_cap[o] * wpm;
    }
    for (int par = 0; par < 2; par++)
        for (uint32_t o = 0; o < world; o++) {
            if (o == target) continue;
            L.header_off[par][o] = off; off += EX_HEADER_BYTES;
            L.len_off[par][o] = off; off = (off + L.m_cap[o] * 4 + 255) & ~255ull;
            L.meta_off[par][o] = off; off = (off + L.m_cap[o] * sizeof(uint4) + 255) & ~255ull;
            L.raw_off[par][o] = off; off = (off + L.m_cap[o] * P * 8 + 255) & ~255ull;
            L.ent_off[par][o] = off; off = (off + L.w_cap[o] * 4 + 255) & ~255ull;
        }
    L.bytes = off;
    return L;
}
static void ex_block(ExBlock &b, uint8_t *base, const ExLayout &L, uint32_t origin) {
    for (int par = 0; par < 2; par++) {
        b.header[par] = (uint64_t *)(base + L.header_off[par][origin]);
        b.len[par] = (uint32_t *)(base + L.len_off[par][origin]);
        b.raw[par] = (double *)(base + L.raw_off[par][origin]);
        b.entries[par] = (uint32_t *)(base + L.ent_off[par][origin]);
        b.flag[par] = (unsigned long long *)(base + L.flag_off[par][origin]);
        b.eflag[par] = (unsigned long long *)(base + L.eflag_off[par][origin]);
        b.meta[par] = (uint4 *)(base + L.meta_off[par][origin]);
    }
    b.m_cap = L.m_cap[origin];
    b.w_cap = L.w_cap[origin];
}
static void ex_free(vl_exchange *ex) {
    if (!ex) return;
    vl_sim *sim = ex->sim;
    cudaSetDevice(sim->device);
    cudaStreamSynchronize(sim->stream);
    for (uint32_t t = 0; t < ex->world; t++) {
        if (ex->peer[t] && ex->peer_ipc[t]) cudaIpcCloseMemHandle(ex->peer[t]);
        if (ex->part_ids[t]) cudaFree(ex->part_ids[t]);
    }
    if (ex->side) { cudaStreamSynchronize(ex->side); cudaStreamDestroy(ex->side); }
    if (ex->ev_drawn) cudaEventDestroy(ex->ev_drawn);
    if (ex->ev_imported) cudaEventDestroy(ex->ev_imported);
    void *bufs[] = {ex->recv, ex->scratch, ex->words_out, ex->words_in, ex->recs_in, ex->rec_off, ex->woff, ex->word_off, ex->stage, ex->own_fit};
    for (void *b : bufs) if (b) cudaFree(b);
    delete ex;
}
// Kernels are loaded lazily, and loading one may wait for every running kernel of the context.  A rank whose stream is
// waiting (on the device) for a peer that lives in the same process must therefore never be the reason a launch blocks:
// every kernel of the exchange loop is loaded before the first generation is enqueued.
static void preload_exchange_kernels() {
    static std::once_flag once;
    std::call_once(once, [] {
        const void *ks[] = {
            (const void *)k_zero_words, (const void *)k_begin_generation_fused, (const void *)k_zero_hsum_current,
            (const void *)k_infect_sum<false>, (const void *)k_infect_sum<true>, (const void *)k_mutate<false>, (const void *)k_mutate<true>, (const void *)k_mutate<false, true, 1>, (const void *)k_mutate<true, true, 1>, (const void *)k_mutate<false, true, 2>, (const void *)k_mutate<true, true, 2>,
            (const void *)k_offspring_sum<false>, (const void *)k_offspring_sum<true>, (const void *)k_plan_resample,
            (const void *)k_resample_pos<false, true>, (const void *)k_resample_pos<true, true>, (const void *)k_resample_pos<true, true, true>, (const void *)k_resample_pos<false, false>, (const void *)k_resample_pos<true, false>, (const void *)k_resample_pos<true, false, true>, (const void *)k_swap_population_fused,
            (const void *)k_ex_sizes, (const void *)k_ex_export,
            (const void *)k_ex_wait, (const void *)k_ex_import_fused, (const void *)k_ex_own_items, (const void *)k_shard_publish_total, (const void *)k_shard_split,
            (const void *)k_ex_out_starts, (const void *)k_ex_lens, (const void *)k_ex_copy, (const void *)k_ex_stream, (const void *)k_ex_publish,
            (const void *)k_ex_in_starts, (const void *)k_ex_in_words, (const void *)k_ex_reserve, (const void *)k_ex_import, (const void *)k_ex_own,
            (const void *)k_swap_population, (const void *)k_scan_u32<uint32_t, true>, (const void *)k_scan_u32<uint64_t, true>,
            (const void *)k_gc_begin, (const void *)k_gc_clear, (const void *)k_gc_mark, (const void *)k_gc_mark_pending, (const void *)k_gc_words,
            (const void *)k_gc_copy, (const void *)k_gc_remap, (const void *)k_gc_remap_pending, (const void *)k_gc_end,
        };
        for (const void *k : ks) {
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, k);
        }
    });
}
VL_API int vl_exchange_create(vl_sim *sim, uint32_t rank, uint32_t world, const double *transfer, uint64_t words_per_migrant,
                              int sharded, uint64_t shard_seed, vl_exchange **out, void *ipc_handle_out) {
    ENTER(sim);
    if (!out || (!transfer && !sharded) || !ipc_handle_out) return fail(sim, VL_ERR_ARGUMENT, "null argument");
    if (world < 2 || world > EX_MAX_WORLD || rank >= world) return fail(sim, VL_ERR_ARGUMENT, "world must be in [2, %u] and rank below it", EX_MAX_WORLD);
    std::lock_guard<std::mutex> context_lock(g_context_mu);
    vl_exchange *ex = new vl_exchange();
    ex->sim = sim; ex->rank = rank; ex->world = world;
    ex->wpm = words_per_migrant ? words_per_migrant : 64;
    ex->sharded = sharded != 0;
    if (transfer) ex->T.assign(transfer, transfer + (size_t)world * world);
    else ex->T.assign((size_t)world * world, 1.0 / (double)world);
    ex->Tcap = ex->T;
    const DevState &h = sim->h;
    ex->max_pop = h.max_population;
    const ExLayout L = ex_layout(world, rank, ex->Tcap.data(), h.max_population, h.n_providers, ex->wpm, ex->sharded);
    ex->recv_bytes = L.bytes;
#define XCU(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { ex_free(ex); return fail(sim, VL_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } } while (0)
    XCU(cudaMalloc((void **)&ex->recv, L.bytes));
    XCU(cudaMemset(ex->recv, 0, L.bytes));
    XCU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)ipc_handle_out, ex->recv));
    XCU(cudaMalloc((void **)&ex->scratch, sizeof(ExScratch)));
    XCU(cudaMemset(ex->scratch, 0, sizeof(ExScratch)));
    memset(&ex->args, 0, sizeof ex->args);
    ex->args.world = world; ex->args.rank = rank; ex->args.n_providers = h.n_providers;
    ex->args.shard_seed = shard_seed;
    for (int par = 0; par < 2; par++) {
        ex->args.tot_in[par] = (uint64_t *)(ex->recv + L.tot_off[par]);
        ex->args.tot_flag_in[par] = (unsigned long long *)(ex->recv + L.tflag_off[par]);
        ex->args.tot_out[rank][par] = ex->args.tot_in[par] + rank;
        ex->args.tot_flag_out[rank][par] = ex->args.tot_flag_in[par] + rank;
    }
    for (uint32_t o = 0; o < world; o++) {
        if (o == rank) continue;
        ex_block(ex->args.in[o], ex->recv, L, o);
        ex->m_in_bound += L.m_cap[o];
    }
    for (uint32_t t = 0; t < world; t++) {  // one id buffer per target part (my column of T)
        double m = ex->sharded ? (double)shard_cell_capacity(h.max_population, world) + 1.0
                               : floor(ex->Tcap[(size_t)t * world + rank] * (double)h.max_population) + 2.0;
        if (!(m >= 2.0)) m = 2.0;
        XCU(cudaMalloc((void **)&ex->part_ids[t], (uint64_t)m * 4));
        ex->args.part[t] = ex->part_ids[t];
        if (t == rank) XCU(cudaMalloc((void **)&ex->own_fit, (uint64_t)m * 8));
        if (t != rank) ex->m_out_bound += (uint64_t)m;
    }
    XCU(cudaMalloc((void **)&ex->words_out, (ex->m_out_bound + 2) * 4));
    XCU(cudaMalloc((void **)&ex->woff, (ex->m_out_bound + 2) * 8));
    XCU(cudaMalloc((void **)&ex->stage, (ex->m_out_bound * ex->wpm + 64) * 4));
    XCU(cudaMalloc((void **)&ex->words_in, (ex->m_in_bound + 2) * 4));
    XCU(cudaMalloc((void **)&ex->recs_in, (ex->m_in_bound + 2) * 4));
    XCU(cudaMalloc((void **)&ex->rec_off, (ex->m_in_bound + 2) * 4));
    XCU(cudaMalloc((void **)&ex->word_off, (ex->m_in_bound + 2) * 8));
    {   // the side stream carries the NVLink-bound kernels: they need few SMs but should get them at once
        int lo = 0, hi = 0;
        XCU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        XCU(cudaStreamCreateWithPriority(&ex->side, cudaStreamNonBlocking, hi));
    }
    XCU(cudaEventCreateWithFlags(&ex->ev_drawn, cudaEventDisableTiming));
    XCU(cudaEventCreateWithFlags(&ex->ev_imported, cudaEventDisableTiming));
#undef XCU
    preload_exchange_kernels();
    {   // the scans of the loop (imports, compaction) must not have to grow their work area (that synchronises)
        const int rc = zero_scan_work(sim, std::max<uint64_t>(h.cap_hap / 32 + 16, std::max(ex->m_in_bound, ex->m_out_bound) + 16));
        if (rc != VL_OK) { ex_free(ex); return rc; }
    }
    *out = ex;
    return VL_OK;
}
VL_API int vl_exchange_recv_buffer(vl_exchange *ex, void **ptr, uint64_t *bytes) {
    if (!ex || !ptr) return fail(nullptr, VL_ERR_ARGUMENT, "null argument");
    *ptr = ex->recv;
    if (bytes) *bytes = ex->recv_bytes;
    return VL_OK;
}
// ipc_handles: world x 64 bytes (cudaIpcMemHandle_t of every rank's receive buffer, own slot ignored); local_ptrs: optional,
// a non-null entry is used as that rank's receive buffer directly (ranks living in this process)
VL_API int vl_exchange_connect(vl_exchange *ex, const void *ipc_handles, void *const *local_ptrs) {
    if (!ex) return fail(nullptr, VL_ERR_ARGUMENT, "null exchange");
    vl_sim *sim = ex->sim;
    ENTER(sim);
    const DevState &h = sim->h;
    for (uint32_t t = 0; t < ex->world; t++) {
        if (t == ex->rank) continue;
        if (local_ptrs && local_ptrs[t]) {
            ex->peer[t] = local_ptrs[t];
        } else {
            if (!ipc_handles) return fail(sim, VL_ERR_ARGUMENT, "no handle for rank %u", t);
            cudaIpcMemHandle_t hd;
            memcpy(&hd, (const char *)ipc_handles + (size_t)t * sizeof(cudaIpcMemHandle_t), sizeof hd);
            CU(sim, cudaIpcOpenMemHandle(&ex->peer[t], hd, cudaIpcMemLazyEnablePeerAccess));
            ex->peer_ipc[t] = true;
        }
        const ExLayout L = ex_layout(ex->world, t, ex->Tcap.data(), h.max_population, h.n_providers, ex->wpm, ex->sharded);
        ex_block(ex->args.out[t], (uint8_t *)ex->peer[t], L, ex->rank);
        for (int par = 0; par < 2; par++) {
            ex->args.tot_out[t][par] = (uint64_t *)((uint8_t *)ex->peer[t] + L.tot_off[par]) + ex->rank;
            ex->args.tot_flag_out[t][par] = (unsigned long long *)((uint8_t *)ex->peer[t] + L.tflag_off[par]) + ex->rank;
        }
    }
    ex->connected = true;
    return VL_OK;
}
// The schedule resolves a transfer matrix per generation (src/config/schedule.rs:196-266, used at src/runner.rs:382-422).
// The buffers keep the sizes of the creation matrix, so every entry of the new one must stay at or below it (create the
// exchange with the entry-wise maximum over the schedule).  Takes effect with the next vl_exchange_push; every rank must
// set the same matrix for the same generation.
VL_API int vl_exchange_set_transfer(vl_exchange *ex, const double *transfer) {
    if (!ex || !transfer) return fail(nullptr, VL_ERR_ARGUMENT, "null argument");
    vl_sim *sim = ex->sim;
    ENTER(sim);
    if (ex->sharded) return fail(sim, VL_ERR_ARGUMENT, "a sharded compartment has no transfer matrix");
    if (ex->pushed) return fail(sim, VL_ERR_IMPLEMENTATION, "a generation of this exchange is in flight: set the matrix after vl_exchange_pull");
    const size_t n = (size_t)ex->world * ex->world;
    for (size_t k = 0; k < n; k++) {
        if (!(transfer[k] >= 0.0) || !(transfer[k] <= ex->Tcap[k]))
            return fail(sim, VL_ERR_ARGUMENT, "transfer[%zu][%zu] = %g is outside [0, %g], the entry the exchange was created with", k / ex->world,
                        k % ex->world, transfer[k], ex->Tcap[k]);
    }
    ex->T.assign(transfer, transfer + n);
    return VL_OK;
}
VL_API int vl_exchange_destroy(vl_exchange *ex) {
    if (!ex) return VL_OK;
    std::lock_guard<std::recursive_mutex> lock_(ex->sim->mu);
    std::lock_guard<std::mutex> context_lock(g_context_mu);
    ex_free(ex);
    return VL_OK;
}
// the generation up to the offspring map; a sharded compartment also publishes its offspring total to every rank
VL_API int vl_exchange_begin(vl_exchange *ex) {
    if (!ex) return fail(nullptr, VL_ERR_ARGUMENT, "null exchange");
    vl_sim *sim = ex->sim;
    ENTER(sim);
    if (!ex->connected) return fail(sim, VL_ERR_IMPLEMENTATION, "vl_exchange_connect has not run");
    if (ex->begun || ex->pushed) return fail(sim, VL_ERR_IMPLEMENTATION, "the previous generation of this exchange is not finished");
    ex->fused_step = fused_now(sim);
    ex->fast_step = ex->fused_step && !(sim->experiment & 256);
    ex->prefill_step = ex->fast_step && ex->more_follow && !(sim->experiment & 32);
    // (measured at 2 x 10^7, K2: the own part's k_mutate and the export slow each other down — both wait on scattered DRAM
    // reads — and the step gets longer, not shorter; the overlapped order stays available for A/B runs)
    ex->overlap_step = ex->prefill_step && (sim->experiment & 1024);
    if (!ex->fused_step) drop_prefill(sim);
    if (ex->fused_step) TRY(enqueue_generation_front_fused(sim));
    else TRY(enqueue_generation_front(sim));
    ex->seq += 1;
    if (ex->sharded) {
        ExArgs a = ex->args;
        a.seq = ex->seq;
        a.parity = (uint32_t)(ex->seq & 1);
        LAUNCH(sim, k_shard_publish_total, 1, TPB, 0, sim->d, a);
    }
    ex->begun = true;
    CU(sim, cudaGetLastError());
    return VL_OK;
}
// the parts of every target and the push into the targets' buffers
VL_API int vl_exchange_push(vl_exchange *ex) {
    if (!ex) return fail(nullptr, VL_ERR_ARGUMENT, "null exchange");
    vl_sim *sim = ex->sim;
    ENTER(sim);
    if (!ex->connected) return fail(sim, VL_ERR_IMPLEMENTATION, "vl_exchange_connect has not run");
    if (ex->pushed) return fail(sim, VL_ERR_IMPLEMENTATION, "vl_exchange_pull must follow vl_exchange_push");
    const uint32_t W = ex->world, r = ex->rank;
    if (!ex->begun) TRY(vl_exchange_begin(ex));
    ex->begun = false;
    ExArgs a = ex->args;
    a.seq = ex->seq;
    a.parity = (uint32_t)(ex->seq & 1);
    double factors[EX_MAX_WORLD];
    uint32_t *dsts[EX_MAX_WORLD];
    for (uint32_t t = 0; t < W; t++) { factors[t] = ex->T[(size_t)t * W + r]; dsts[t] = ex->part_ids[t]; }
    double *fits[EX_MAX_WORLD];
    for (uint32_t t = 0; t < W; t++) fits[t] = t == r ? ex->own_fit : nullptr;  // (migrants get theirs from the record they become)
    a.own_fit = ex->fused_step ? ex->own_fit : nullptr;
    ExScratch *sc = ex->scratch;
    PosOpts po;
    if (ex->fast_step) {
        // the part sizes are final once the plan has run: announce them, learn every origin's, and draw my own part
        // straight into its place in the new population (with the children's next infection when another generation follows)
        po.own_part = (int)r;
        po.own_off = (const uint64_t *)((const char *)sc + offsetof(ExScratch, own_off));
        po.prefill = ex->prefill_step;
        po.predict_ids = false;  // the imported migrants become records before the next k_mutate runs
        po.between = [=]() -> int {
            LAUNCH(sim, k_ex_sizes, 1, 32, 0, sim->d, a, sc, 20ull * 1000000000ull);
            return VL_OK;
        };
    }
    if (ex->sharded) {
        // every rank's offspring total has been published (vl_exchange_begin): wait for them, split the N' draws over
        // the (target, origin) cells, draw my row of parts with the amounts the split left in DevState::plan_amount
        LAUNCH(sim, k_shard_split, 1, 32, 0, sim->d, a, 20ull * 1000000000ull);
        TRY(launch_plan_resample(sim, W, nullptr, 2, nullptr, 0, dsts, nullptr, 0, ex->fused_step, fits, &po));
    } else {
        TRY(launch_plan_resample(sim, W, factors, 0, nullptr, 0, dsts, nullptr, 0, ex->fused_step, fits, &po));
    }
    if (ex->fast_step) {
        const unsigned eg = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (ex->m_out_bound + EXP_TPB - 1) / EXP_TPB), (uint64_t)sim->n_sms * 4);
        if (ex->overlap_step) {
            static const int side_ctas = getenv("VL_EX_CTAS") ? atoi(getenv("VL_EX_CTAS")) : 2;
            const unsigned eg = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (ex->m_out_bound + EXP_TPB - 1) / EXP_TPB), (uint64_t)sim->n_sms * side_ctas);
            // The migrants leave on the side stream.  Meanwhile this stream mutates the children just drawn for the next
            // generation (their worklist items and the ids of their new records are fixed first: the import appends
            // items and creates records of its own while that runs).
            LAUNCH(sim, k_ex_own_items, 1, 1, 0, sim->d);
            CU(sim, cudaEventRecord(ex->ev_drawn, sim->stream));
            CU(sim, cudaStreamWaitEvent(ex->side, ex->ev_drawn, 0));
            {
                StreamScope on_side(sim, ex->side);
                LAUNCH(sim, k_ex_export, eg, EXP_TPB, EXP_SMEM_BYTES, sim->d, a, sc, (sim->experiment & 512) ? 0 : 1);
            }
            LAUNCH_MUTATE_FUSED(sim, grid_for(sim, sim->n_bound / 4 + 1, 8), TPB, mutate_smem_bytes(sim->h.L), sim->d, ReplayMut{nullptr, nullptr, nullptr}, 0, 4, 1, 1);
        } else {
            LAUNCH(sim, k_ex_export, eg, EXP_TPB, EXP_SMEM_BYTES, sim->d, a, sc, (sim->experiment & 512) ? 0 : 1);
        }
        ex->pushed = true;
        CU(sim, cudaGetLastError());
        return VL_OK;
    }
    const uint64_t *n_out = (const uint64_t *)((const char *)sc + offsetof(ExScratch, n_out));
    // ---- push (general kernels: any flow, several passes)
    LAUNCH(sim, k_ex_out_starts, 1, 1, 0, sim->d, a, sc);
    LAUNCH(sim, k_ex_lens, grid_for(sim, ex->m_out_bound), TPB, 0, sim->d, a, (const ExScratch *)sc, ex->words_out);
    TRY((launch_scan<uint64_t, true>(sim, ex->words_out, ex->woff, n_out, 0, ex->m_out_bound, nullptr, 1, nullptr)));
    LAUNCH(sim, k_ex_copy, grid_for(sim, ex->m_out_bound), TPB, 0, sim->d, a, (const ExScratch *)sc, (const uint64_t *)ex->woff, ex->stage, ex->m_out_bound * ex->wpm);
    LAUNCH(sim, k_ex_stream, grid_for(sim, ex->m_out_bound * 4), TPB, 0, sim->d, a, (const ExScratch *)sc, (const uint64_t *)ex->woff, (const uint32_t *)ex->stage,
           ex->m_out_bound * ex->wpm);
    LAUNCH(sim, k_ex_publish, 1, 32, 0, sim->d, a, (const ExScratch *)sc, (const uint64_t *)ex->woff);
    ex->pushed = true;
    CU(sim, cudaGetLastError());
    return VL_OK;
}
// second half: wait (on the device) for every origin's block of this generation, import, install the new population
VL_API int vl_exchange_pull(vl_exchange *ex) {
    if (!ex) return fail(nullptr, VL_ERR_ARGUMENT, "null exchange");
    vl_sim *sim = ex->sim;
    ENTER(sim);
    if (!ex->pushed) return fail(sim, VL_ERR_IMPLEMENTATION, "vl_exchange_push has not run for this generation");
    ex->pushed = false;
    const uint32_t W = ex->world;
    ExArgs a = ex->args;
    a.seq = ex->seq;
    a.parity = (uint32_t)(ex->seq & 1);
    ExScratch *sc = ex->scratch;
    const uint64_t *n_imp = (const uint64_t *)((const char *)sc + offsetof(ExScratch, n_imp));
    uint64_t *totals = (uint64_t *)((char *)sc + offsetof(ExScratch, totals));
    if (ex->fast_step) {
        const unsigned ig = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (ex->m_in_bound + EXP_TPB - 1) / EXP_TPB), (uint64_t)sim->n_sms * 4);
        {
            StreamScope on_side(sim, ex->overlap_step ? ex->side : sim->stream);
            LAUNCH(sim, k_ex_import_fused, ig, EXP_TPB, 0, sim->d, a, (const ExScratch *)sc, ex->prefill_step ? 1 : 0, 20ull * 1000000000ull);
        }
        if (ex->overlap_step) {
            CU(sim, cudaEventRecord(ex->ev_imported, ex->side));
            CU(sim, cudaStreamWaitEvent(sim->stream, ex->ev_imported, 0));
        }
    } else {
    LAUNCH(sim, k_ex_wait, 1, 32, 0, sim->d, a, 20ull * 1000000000ull);
    LAUNCH(sim, k_ex_in_starts, 1, 1, 0, sim->d, a, sc);
    LAUNCH(sim, k_ex_in_words, grid_for(sim, ex->m_in_bound), TPB, 0, a, (const ExScratch *)sc, ex->words_in, ex->recs_in);
    TRY((launch_scan<uint32_t, true>(sim, ex->recs_in, ex->rec_off, n_imp, 0, ex->m_in_bound, totals + 0, 0, nullptr)));
    TRY((launch_scan<uint64_t, true>(sim, ex->words_in, ex->word_off, n_imp, 0, ex->m_in_bound, totals + 1, 1, nullptr)));
    LAUNCH(sim, k_ex_reserve, 1, 1, 0, sim->d, sc);
    LAUNCH(sim, k_ex_import, grid_for(sim, ex->m_in_bound), TPB, 0, sim->d, a, (const ExScratch *)sc, (const uint32_t *)ex->rec_off, (const uint64_t *)ex->word_off);
    a.own_fit = ex->fused_step ? ex->own_fit : nullptr;
    LAUNCH(sim, k_ex_own, grid_for(sim, sim->h.max_population), TPB, 0, sim->d, a, (const ExScratch *)sc);
    }
    if (ex->fused_step) {  // both per-infectant arrays swap, the per-host sums change sides (k_offspring_sum cleared the other buffer)
        LAUNCH(sim, k_swap_population_fused, 1, 1, 0, sim->d);
        std::swap(sim->h.hap_id, sim->h.hap_id_next);
        std::swap(sim->h.pop_fit, sim->h.pop_fit_next);
    } else {
        swap_population(sim);
    }
    sim->pop_fit_valid = ex->fused_step;  // k_ex_import and k_ex_own filled the carried fitness of the new population
    sim->prefilled = ex->prefill_step;
    sim->prefill_patch = ex->prefill_step;
    sim->prefill_own_done = ex->overlap_step;
    sim->n_valid = false;
    {   // upper bound of the new population: every origin sends at most its block capacity, the own part its cap
        uint64_t bound = 0;
        for (uint32_t o = 0; o < W; o++) {
            double m = ex->sharded ? (double)shard_cell_capacity(ex->max_pop, W) + 1.0
                                   : floor(ex->T[(size_t)ex->rank * W + o] * (double)sim->h.max_population) + 1.0;
            bound += (uint64_t)(m >= 1.0 ? m : 1.0);
        }
        sim->n_bound = std::min<uint64_t>(sim->h.cap_inf, bound);
    }
    sim->csr_valid = false;
    sim->slices_ordered = false;
    sim->layout_valid = false;
    CU(sim, cudaGetLastError());
    return maybe_gc(sim);
}
// one generation of Runner::run's loop body (src/runner.rs:382-422) for this rank's compartment, transmission included;
// everything is enqueued, nothing is read back
VL_API int vl_exchange_step(vl_exchange *ex) {
    TRY(vl_exchange_push(ex));
    return vl_exchange_pull(ex);
}
// n generations enqueued back to back (the ranks meet on the device only).  Every generation but the last also prepares the
// infection of the next one while it draws and imports (as vl_run does); nothing of that prepared state outlives the call.
VL_API int vl_exchange_run(vl_exchange *ex, uint64_t n_generations) {
    if (!ex) return fail(nullptr, VL_ERR_ARGUMENT, "null exchange");
    std::lock_guard<std::recursive_mutex> lock_(ex->sim->mu);
    for (uint64_t g = 0; g < n_generations; g++) {
        ex->more_follow = g + 1 < n_generations;
        int rc = vl_exchange_push(ex);
        if (rc == VL_OK) rc = vl_exchange_pull(ex);
        ex->more_follow = false;
        if (rc != VL_OK) return rc;
    }
    return VL_OK;
}

static inline uint64_t f64_as_usize_host(double x) {
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551615.0) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)x;
}

VL_API int vl_step(vl_sim *const *sims, uint32_t C, const double *T, int rule) {
    if (!sims || C == 0 || !T) return fail(nullptr, VL_ERR_ARGUMENT, "null argument");
    std::vector<std::unique_lock<std::recursive_mutex>> locks;
    for (uint32_t c = 0; c < C; c++) {
        if (!sims[c]) return fail(nullptr, VL_ERR_ARGUMENT, "null handle");
        locks.emplace_back(sims[c]->mu);
    }
    vl_sim *s0 = sims[0];
#define ON(c) CU(sims[c], cudaSetDevice(sims[c]->device))
    // runner.rs:382-399: generation++, infect, mutate, replicate on every compartment (streams overlap)
    for (uint32_t c = 0; c < C; c++) {
        ON(c);
        LAUNCH(sims[c], k_tick, 1, 1, 0, sims[c]->d);
        sims[c]->salt = 0;
        TRY(enqueue_infect(sims[c]));
        TRY(enqueue_mutate(sims[c]));
        TRY(enqueue_replicate(sims[c]));
    }
    std::vector<vl_pop *> parts((size_t)C * C, nullptr);
    auto cleanup = [&]() { for (vl_pop *p : parts) if (p) { cudaSetDevice(p->owner->device); cudaStreamSynchronize(p->owner->stream); drop_pop(p); } };
    int rc = VL_OK;
    if (rule == 0) {
        // parallel-build rule (runner.rs:401-416): new[t] = concat_o subsample(sim[o], offspring[o], T[t][o])
        std::vector<double> col(C);
        std::vector<vl_pop *> out(C);
        for (uint32_t o = 0; o < C && rc == VL_OK; o++) {  // the C parts of one origin share one plan + resample pass
            ON(o);
            for (uint32_t t = 0; t < C; t++) col[t] = T[(size_t)t * C + o];
            rc = enqueue_subsamples(sims[o], C, col.data(), 0, nullptr, out.data());
            for (uint32_t t = 0; t < C; t++) parts[(size_t)t * C + o] = out[t];
        }
    } else {
        // serial-build rule (runner.rs:550-631)
        std::vector<double> ts(C);
        for (uint32_t t = 0; t < C && rc == VL_OK; t++) {
            ON(t);
            LAUNCH(sims[t], k_target_size, 1, TPB, 0, sims[t]->d);
            rc = read_field(sims[t], FIELD_OFF(target_size), &ts[t]);
        }
        std::vector<uint64_t> amount((size_t)C * C, 0);
        if (rc == VL_OK) {
            uint32_t gen = 0;
            rc = read_field(s0, FIELD_OFF(generation), &gen);
            for (uint32_t t = 0; t < C; t++) {
                Philox g(sims[t]->h.seed, sims[t]->h.compartment, gen, PH_MIGRANT, t);
                for (uint32_t o = 0; o < C; o++) {
                    if (t == o) continue;
                    double m = ts[t] * T[(size_t)t * C + o];
                    uint64_t residue = g.u01() < m - floor(m) ? 1 : 0;
                    amount[(size_t)t * C + o] = f64_as_usize_host(m) + residue;
                }
            }
            for (uint32_t t = 0; t < C && rc == VL_OK; t++) {
                uint64_t mig = 0;
                for (uint32_t o = 0; o < C; o++) mig += amount[(size_t)t * C + o];
                if (f64_as_usize_host(ts[t]) < mig) rc = fail(s0, VL_ERR_REFERENCE_PANIC, "migrants into compartment %u exceed its target size: usize underflow panics (src/runner.rs:583-590)", t);
                else amount[(size_t)t * C + t] = f64_as_usize_host(ts[t]) - mig;
            }
        }
        std::vector<uint64_t> col(C);
        std::vector<vl_pop *> out(C);
        for (uint32_t o = 0; o < C && rc == VL_OK; o++) {
            ON(o);
            for (uint32_t t = 0; t < C; t++) col[t] = amount[(size_t)t * C + o];
            rc = enqueue_subsamples(sims[o], C, nullptr, 1, col.data(), out.data());
            for (uint32_t t = 0; t < C; t++) parts[(size_t)t * C + o] = out[t];
        }
    }
    for (uint32_t c = 0; c < C && rc == VL_OK; c++) {
        ON(c);
        cudaStreamSynchronize(sims[c]->stream);
        rc = check_device_error(sims[c]);
        if (rc != VL_OK && c != 0) fail(s0, rc, "%s", sims[c]->err);
    }
    if (rc != VL_OK) { cleanup(); return rc; }
    for (vl_pop *p : parts) finalize_pop(p);
    // every compartment's new population must fit before ANY is installed: a step either happens for all compartments
    // or (for this foreseeable failure) for none
    for (uint32_t t = 0; t < C; t++) {
        uint64_t total = 0;
        for (uint32_t o = 0; o < C; o++) total += parts[(size_t)t * C + o] ? parts[(size_t)t * C + o]->n : 0;
        if (total > sims[t]->h.cap_inf) {
            cleanup();
            return fail(s0, VL_ERR_CAPACITY, "compartment %u would receive %llu infectants, its capacity_infectants is %llu; no population was replaced", t,
                        (unsigned long long)total, (unsigned long long)sims[t]->h.cap_inf);
        }
    }
    for (uint32_t t = 0; t < C; t++) {
        ON(t);
        rc = vl_set_population(sims[t], &parts[(size_t)t * C], C);
        if (rc != VL_OK) {  // (a CUDA or record-capacity failure half way: compartments below t hold the new generation, the others the old one)
            fail(s0, rc, "installing compartment %u failed, the handles of this step are no longer consistent: %s", t, sims[t]->err);
            cleanup();
            return rc;
        }
        for (uint32_t o = 0; o < C; o++) parts[(size_t)t * C + o] = nullptr;  // consumed by vl_set_population
    }
#undef ON
    return VL_OK;
}

VL_API int vl_infectant_generations(vl_sim *sim, uint64_t *out, int reset) {
    ENTER(sim);
    TRY(read_field(sim, FIELD_OFF(infectant_generations), out));
    if (reset) TRY(write_field(sim, FIELD_OFF(infectant_generations), (uint64_t)0));
    return VL_OK;
}

// ---------------------------------------------------------------------------------------------- replay hooks
template <typename T>
static int upload_replace(vl_sim *sim, T **dev, const T *src, uint64_t n) {
    CU(sim, cudaStreamSynchronize(sim->stream));
    if (*dev) { cudaFree(*dev); *dev = nullptr; }
    CU(sim, cudaMalloc((void **)dev, std::max<uint64_t>(n, 1) * sizeof(T)));
    if (n) CU(sim, cudaMemcpy(*dev, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return VL_OK;
}
VL_API int vl_replay_infections(vl_sim *sim, const int64_t *infections, uint64_t n) {
    ENTER(sim);
    drop_prefill(sim);
    uint64_t cur = 0;
    TRY(population_len(sim, &cur));
    if (n != cur) return fail(sim, VL_ERR_ARGUMENT, "infection log has %llu entries, population has %llu", (unsigned long long)n, (unsigned long long)cur);
    TRY(upload_replace(sim, &sim->rp_inf, infections, n));
    sim->has_rp_inf = true;
    return VL_OK;
}
VL_API int vl_replay_mutations(vl_sim *sim, const uint64_t *offsets, uint64_t n, const uint32_t *sites, const uint8_t *bases) {
    ENTER(sim);
    uint64_t cur = 0;
    TRY(population_len(sim, &cur));
    if (n != cur) return fail(sim, VL_ERR_ARGUMENT, "mutation log has %llu infectants, population has %llu", (unsigned long long)n, (unsigned long long)cur);
    for (uint64_t i = 0; i < n; i++) if (offsets[i + 1] < offsets[i]) return fail(sim, VL_ERR_ARGUMENT, "mutation log offsets not monotone");
    const uint64_t total = n ? offsets[n] : 0;
    uint64_t zero = 0;
    TRY(upload_replace(sim, &sim->rp_mut_off, n ? offsets : &zero, n + 1));
    TRY(upload_replace(sim, &sim->rp_mut_sites, sites, total));
    TRY(upload_replace(sim, &sim->rp_mut_bases, bases, total));
    sim->has_rp_mut = true;
    return VL_OK;
}
VL_API int vl_replay_recombinations(vl_sim *sim, uint64_t n_events, const uint64_t *host, const uint32_t *i, const uint32_t *j,
                                    const uint32_t *s0, const uint32_t *s1, const uint8_t *which) {
    ENTER(sim);
    for (uint64_t e = 1; e < n_events; e++) if (host[e] < host[e - 1]) return fail(sim, VL_ERR_ARGUMENT, "recombination log must be sorted by host");
    TRY(upload_replace(sim, &sim->rp_rec_host, host, n_events));
    TRY(upload_replace(sim, &sim->rp_rec_i, i, n_events));
    TRY(upload_replace(sim, &sim->rp_rec_j, j, n_events));
    TRY(upload_replace(sim, &sim->rp_rec_s0, s0, n_events));
    TRY(upload_replace(sim, &sim->rp_rec_s1, s1, n_events));
    TRY(upload_replace(sim, &sim->rp_rec_which, which, n_events));
    sim->rp_rec_n = n_events;
    sim->has_rp_rec = true;
    return VL_OK;
}
VL_API int vl_replay_offspring(vl_sim *sim, const uint64_t *offspring, uint64_t n) {
    ENTER(sim);
    drop_prefill(sim);
    uint64_t cur = 0;
    TRY(population_len(sim, &cur));
    if (n != cur) return fail(sim, VL_ERR_ARGUMENT, "offspring log has %llu entries, population has %llu", (unsigned long long)n, (unsigned long long)cur);
    TRY(upload_replace(sim, &sim->rp_off, offspring, n));
    sim->has_rp_off = true;
    return VL_OK;
}

// ---------------------------------------------------------------------------------------------- inspection
template <typename T>
static int download(vl_sim *sim, std::vector<T> &host, const T *dev, uint64_t n) {
    host.resize(n);
    if (n) CU(sim, cudaMemcpyAsync(host.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    return VL_OK;
}
static int expect_len(vl_sim *sim, uint64_t n) {
    uint64_t cur = 0;
    TRY(population_len(sim, &cur));
    if (n != cur) return fail(sim, VL_ERR_ARGUMENT, "buffer has %llu entries, population has %llu", (unsigned long long)n, (unsigned long long)cur);
    return VL_OK;
}
VL_API int vl_get_infections(vl_sim *sim, int64_t *out, uint64_t n) {
    ENTER(sim);
    TRY(expect_len(sim, n));
    std::vector<uint32_t> v;
    TRY(download(sim, v, (const uint32_t *)sim->h.host_id, n));
    for (uint64_t i = 0; i < n; i++) out[i] = v[i] == NONE32 ? -1 : (int64_t)v[i];
    return VL_OK;
}
VL_API int vl_get_host_map(vl_sim *sim, uint64_t *offsets_out, uint64_t *hosts_out, uint64_t *n_infected_out) {
    ENTER(sim);
    TRY(ensure_ordered_slices(sim));
    std::vector<uint32_t> off, hosts;
    TRY(download(sim, off, (const uint32_t *)sim->h.offsets, sim->h.H + 1));
    const uint64_t ninf = off[sim->h.H];
    uint32_t *tmp = sim->h.rank;  // per-infectant scratch, free once the records are scattered
    LAUNCH(sim, k_extract_hosts, grid_for(sim, ninf), TPB, 0, sim->d, tmp);
    TRY(download(sim, hosts, (const uint32_t *)tmp, ninf));
    if (offsets_out) for (uint64_t k = 0; k <= sim->h.H; k++) offsets_out[k] = off[k];
    if (hosts_out) for (uint64_t k = 0; k < ninf; k++) hosts_out[k] = hosts[k];
    if (n_infected_out) *n_infected_out = ninf;
    return check_device_error(sim);
}
VL_API int vl_get_replication_terms(vl_sim *sim, double *fitness_out, double *host_sum_out, double *lambda_out, uint64_t n) {
    ENTER(sim);
    TRY(expect_len(sim, n));
    if (!sim->csr_valid || !sim->layout_valid) return fail(sim, VL_ERR_IMPLEMENTATION, "no replication terms: vl_replicate_infectants has not run on this population");
    double *tmp = nullptr;
    CU(sim, cudaMalloc((void **)&tmp, std::max<uint64_t>(3 * n, 1) * 8));
    CU(sim, cudaMemsetAsync(tmp, 0xFF, std::max<uint64_t>(3 * n, 1) * 8, sim->stream));  // NaN where not infected
    LAUNCH(sim, k_replication_terms, grid_for(sim, sim->h.H), TPB, 0, sim->d, tmp, tmp + n, tmp + 2 * n);
    if (n) {
        if (fitness_out) CU(sim, cudaMemcpyAsync(fitness_out, tmp, n * 8, cudaMemcpyDeviceToHost, sim->stream));
        if (host_sum_out) CU(sim, cudaMemcpyAsync(host_sum_out, tmp + n, n * 8, cudaMemcpyDeviceToHost, sim->stream));
        if (lambda_out) CU(sim, cudaMemcpyAsync(lambda_out, tmp + 2 * n, n * 8, cudaMemcpyDeviceToHost, sim->stream));
    }
    CU(sim, cudaStreamSynchronize(sim->stream));
    cudaFree(tmp);
    return VL_OK;
}
VL_API int vl_get_offspring_prefix(vl_sim *sim, uint64_t *prefix_out, uint64_t n) {
    ENTER(sim);
    TRY(expect_len(sim, n));
    if (!sim->layout_valid) return fail(sim, VL_ERR_IMPLEMENTATION, "no offspring map on the device");
    // inclusive scan of the offspring map in population order (the CDF Population::sample draws from)
    uint32_t *tmp = sim->h.rank;
    if (n) CU(sim, cudaMemsetAsync(tmp, 0, n * 4, sim->stream));
    LAUNCH(sim, k_unsort_offspring32, grid_for(sim, n), TPB, 0, sim->d, tmp);
    TRY((launch_scan<uint64_t, false>(sim, tmp, sim->stage64, nullptr, n, n, nullptr, 0, nullptr)));
    if (n) CU(sim, cudaMemcpyAsync(prefix_out, sim->stage64, n * 8, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    return VL_OK;
}
VL_API int vl_get_raw_fitness(vl_sim *sim, uint32_t spec, int which, double *out, uint64_t n) {
    ENTER(sim);
    TRY(expect_len(sim, n));
    if (spec >= sim->h.n_specs) return fail(sim, VL_ERR_ARGUMENT, "spec out of range");
    const int pi = which ? sim->h.specs[spec].infect : sim->h.specs[spec].repro;
    if (pi < 0) { for (uint64_t i = 0; i < n; i++) out[i] = 1.0; return VL_OK; }
    double *tmp = (double *)sim->stage64;
    LAUNCH(sim, k_gather_raw, grid_for(sim, n), TPB, 0, sim->d, (uint32_t)pi, tmp);
    if (n) CU(sim, cudaMemcpyAsync(out, tmp, n * 8, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    return check_device_error(sim);
}
VL_API int vl_export_mutations(vl_sim *sim, uint64_t *offsets_out, uint64_t n, uint32_t *positions_out, uint8_t *bases_out, uint64_t *n_entries) {
    ENTER(sim);
    TRY(expect_len(sim, n));
    uint32_t *cnt = sim->h.rank;  // per-infectant scratch (the mutation worklist may hold the next generation's items: k_resample_pos<true>)
    LAUNCH(sim, k_export_count, grid_for(sim, n), TPB, 0, sim->d, cnt);
    TRY((launch_scan<uint64_t, true>(sim, cnt, sim->stage64, nullptr, n, n, sim->gc_totals + 6, 0, nullptr)));
    CU(sim, cudaMemcpyAsync(sim->pin, sim->gc_totals + 6, 8, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    const uint64_t total = sim->pin[0];
    TRY(check_device_error(sim));  // (whatever was enqueued before has run: a population that a failed kernel left behind is not exported)
    if (n_entries) *n_entries = total;
    if (offsets_out) {
        if (n) CU(sim, cudaMemcpy(offsets_out, sim->stage64, n * 8, cudaMemcpyDeviceToHost));
        offsets_out[n] = total;
    }
    if (positions_out && bases_out && total) {
        uint32_t *dp = nullptr; uint8_t *db = nullptr;
        CU(sim, cudaMalloc((void **)&dp, total * 4));
        CU(sim, cudaMalloc((void **)&db, total));
        LAUNCH(sim, k_export_fill, grid_for(sim, n), TPB, 0, sim->d, (const uint64_t *)sim->stage64, dp, db);
        CU(sim, cudaMemcpyAsync(positions_out, dp, total * 4, cudaMemcpyDeviceToHost, sim->stream));
        CU(sim, cudaMemcpyAsync(bases_out, db, total, cudaMemcpyDeviceToHost, sim->stream));
        CU(sim, cudaStreamSynchronize(sim->stream));
        cudaFree(dp); cudaFree(db);
    }
    return VL_OK;
}
// PopulationFrequencies::frequencies (src/stats/population.rs:15-40): L x 4 relative frequencies, row-major
VL_API int vl_stats_frequencies(vl_sim *sim, double *out, uint64_t n_sites) {
    ENTER(sim);
    const uint64_t L = sim->h.L;
    if (!out || n_sites != L) return fail(sim, VL_ERR_ARGUMENT, "frequency buffer must hold n_sites x 4 values");
    uint64_t n = 0;
    TRY(population_len(sim, &n));
    uint32_t *cnt = nullptr;
    CU(sim, cudaMalloc((void **)&cnt, L * 4 * 4));
    zero_device(sim, cnt, L * 4 * 4);
    LAUNCH(sim, k_stats_counts, grid_for(sim, n), TPB, 0, sim->d, cnt);
    std::vector<uint32_t> c(L * 4);
    std::vector<uint8_t> wt(L);
    cudaError_t e1 = cudaMemcpyAsync(c.data(), cnt, L * 16, cudaMemcpyDeviceToHost, sim->stream);
    cudaError_t e2 = cudaMemcpyAsync(wt.data(), sim->h.wildtype, L, cudaMemcpyDeviceToHost, sim->stream);
    cudaError_t e3 = cudaStreamSynchronize(sim->stream);
    cudaFree(cnt);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) return fail(sim, VL_ERR_CUDA, "frequency download failed");
    for (uint64_t pos = 0; pos < L; pos++) {
        // counts[wildtype] = population_size - sum of the row (the reference overwrites the slot, :28-33)
        int64_t row = (int64_t)c[pos * 4] + c[pos * 4 + 1] + c[pos * 4 + 2] + c[pos * 4 + 3];
        int64_t v[4] = {c[pos * 4], c[pos * 4 + 1], c[pos * 4 + 2], c[pos * 4 + 3]};
        v[wt[pos]] = (int64_t)n - row;
        for (int b = 0; b < 4; b++) out[pos * 4 + b] = (double)v[b] / (double)n;
    }
    return VL_OK;
}
// PopulationStatistics::mean (src/stats/population.rs:50-56) of the mapped fitness under provider (spec, which)
VL_API int vl_stats_mean_fitness(vl_sim *sim, uint32_t spec, int which, double *out) {
    ENTER(sim);
    if (!out) return fail(sim, VL_ERR_ARGUMENT, "null out");
    if (spec >= sim->h.n_specs) return fail(sim, VL_ERR_ARGUMENT, "spec out of range");
    const int pi = which ? sim->h.specs[spec].infect : sim->h.specs[spec].repro;
    uint64_t n = 0;
    TRY(population_len(sim, &n));
    const unsigned g = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (n + 4 * TPB - 1) / (4 * TPB)), 1024);
    double *partial = (double *)sim->stage64;  // >= 4096 doubles
    LAUNCH(sim, k_stats_fitness_sum, g, TPB, 0, sim->d, pi, partial);
    std::vector<double> p(g);
    CU(sim, cudaMemcpyAsync(p.data(), partial, g * 8, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    double sum = 0.0;
    for (unsigned k = 0; k < g; k++) sum += p[k];
    *out = sum / (double)n;  // NaN for an empty population, like 0.0 / 0 in the reference
    return VL_OK;
}
VL_API int vl_get_base(vl_sim *sim, uint64_t infectant, uint64_t position, uint8_t *out) {
    ENTER(sim);
    uint32_t *tmp = (uint32_t *)(sim->gc_totals + 8);
    LAUNCH(sim, k_get_base, 1, 1, 0, sim->d, infectant, (uint32_t)position, tmp);
    CU(sim, cudaMemcpyAsync(sim->pin, tmp, 4, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    uint32_t v;
    memcpy(&v, sim->pin, 4);
    if (v > 3) return fail(sim, VL_ERR_ARGUMENT, "infectant or position out of range");
    *out = (uint8_t)v;
    return VL_OK;
}
VL_API int vl_get_haplotype_ids(vl_sim *sim, uint32_t *out, uint64_t n) {
    ENTER(sim);
    TRY(expect_len(sim, n));
    if (n) CU(sim, cudaMemcpyAsync(out, sim->h.hap_id, n * 4, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    return VL_OK;
}
VL_API int vl_get_counters(vl_sim *sim, uint64_t *out4) {
    ENTER(sim);
    unsigned long long nh = 0, at = 0;
    uint64_t runs = 0;
    TRY(read_field(sim, FIELD_OFF(n_hap), &nh));
    TRY(read_field(sim, FIELD_OFF(arena_top), &at));
    TRY(read_field(sim, FIELD_OFF(gc_runs), &runs));
    out4[0] = nh; out4[1] = at; out4[2] = runs; out4[3] = sim->launches;
    return VL_OK;
}
VL_API int vl_get_capacities(vl_sim *sim, uint64_t *out4) {
    ENTER(sim);
    out4[0] = sim->h.cap_inf; out4[1] = sim->h.cap_hap; out4[2] = sim->h.cap_arena; out4[3] = sim->gc_period;
    return VL_OK;
}
__global__ void __launch_bounds__(TPB) k_debug_last_generation(DevState *st, uint64_t n, uint64_t *off, int64_t *hosts, double *sums, double *fit) {
    // the generation has swapped the population: what it started with sits in the *_next buffers, its sums in the other hsum
    const double *hs = st->hsum[st->hsum_parity ^ 1u];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t h = st->host_id[i];
        off[i] = st->offspring[i];
        hosts[i] = h;
        sums[i] = hs[h];
        fit[i] = st->pop_fit_next[i];
    }
}
VL_API int vl_debug_last_generation(vl_sim *sim, uint64_t *offspring_out, int64_t *hosts_out, double *host_sums_out, double *fitness_out, uint64_t n) {
    ENTER(sim);
    if (!offspring_out || !hosts_out || !host_sums_out || !fitness_out) return fail(sim, VL_ERR_ARGUMENT, "null argument");
    if (!sim->h.hsum[0] || n > sim->h.cap_inf) return fail(sim, VL_ERR_IMPLEMENTATION, "no device-resident generation has run on this handle");
    void *tmp = nullptr;
    CU(sim, cudaMalloc(&tmp, std::max<uint64_t>(n, 1) * 32));
    uint64_t *d_off = (uint64_t *)tmp;
    int64_t *d_h = (int64_t *)(d_off + n);
    double *d_s = (double *)(d_h + n), *d_f = d_s + n;
    LAUNCH(sim, k_debug_last_generation, grid_for(sim, n), TPB, 0, sim->d, n, d_off, d_h, d_s, d_f);
    cudaError_t e = cudaSuccess;
    if (n) {
        e = cudaMemcpyAsync(offspring_out, d_off, n * 8, cudaMemcpyDeviceToHost, sim->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hosts_out, d_h, n * 8, cudaMemcpyDeviceToHost, sim->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(host_sums_out, d_s, n * 8, cudaMemcpyDeviceToHost, sim->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(fitness_out, d_f, n * 8, cudaMemcpyDeviceToHost, sim->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(sim->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(sim, VL_ERR_CUDA, "download failed: %s", cudaGetErrorString(e));
    return VL_OK;
}
VL_API int vl_debug_stamps(vl_sim *sim, uint64_t *out8) {
    ENTER(sim);
    CU(sim, cudaMemcpyAsync(sim->pin + 16, (const char *)sim->d + FIELD_OFF(dbg), 64, cudaMemcpyDeviceToHost, sim->stream));
    CU(sim, cudaStreamSynchronize(sim->stream));
    memcpy(out8, sim->pin + 16, 64);
    return VL_OK;
}
VL_API int vl_collect_garbage(vl_sim *sim) {
    ENTER(sim);
    TRY(enqueue_gc(sim, 1));
    CU(sim, cudaStreamSynchronize(sim->stream));
    return check_device_error(sim);
}

// ---------------------------------------------------------------------------------------------- measurement hooks
VL_API int vl_timer_start(vl_sim *sim) {
    ENTER(sim);
    if (!sim->timer_a) { CU(sim, cudaEventCreate(&sim->timer_a)); CU(sim, cudaEventCreate(&sim->timer_b)); }
    CU(sim, cudaEventRecord(sim->timer_a, sim->stream));
    return VL_OK;
}
VL_API int vl_timer_stop(vl_sim *sim, double *ms_out) {
    ENTER(sim);
    if (!sim->timer_a) return fail(sim, VL_ERR_ARGUMENT, "timer not started");
    CU(sim, cudaEventRecord(sim->timer_b, sim->stream));
    CU(sim, cudaEventSynchronize(sim->timer_b));
    float ms = 0.f;
    CU(sim, cudaEventElapsedTime(&ms, sim->timer_a, sim->timer_b));
    if (ms_out) *ms_out = ms;
    return check_device_error(sim);
}
VL_API int vl_profile(vl_sim *sim, int enable) {
    ENTER(sim);
    CU(sim, cudaStreamSynchronize(sim->stream));
    for (auto &r : sim->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    sim->prof.clear();
    sim->profiling = enable != 0;
    return VL_OK;
}
// text report, one line per kernel: "<name> <launches> <total_ms>\n"
VL_API int vl_profile_report(vl_sim *sim, char *buf, uint64_t cap) {
    ENTER(sim);
    CU(sim, cudaStreamSynchronize(sim->stream));
    std::vector<std::string> names;
    std::vector<double> total;
    std::vector<uint64_t> count;
    for (auto &r : sim->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
        size_t k = 0;
        for (; k < names.size(); k++) if (names[k] == r.name) break;
        if (k == names.size()) { names.push_back(r.name); total.push_back(0.0); count.push_back(0); }
        total[k] += ms;
        count[k] += 1;
    }
    std::string out;
    char line[256];
    for (size_t k = 0; k < names.size(); k++) {
        snprintf(line, sizeof line, "%s %llu %.6f\n", names[k].c_str(), (unsigned long long)count[k], total[k]);
        out += line;
    }
    if (buf && cap) {
        size_t m = std::min<size_t>(out.size(), cap - 1);
        memcpy(buf, out.data(), m);
        buf[m] = 0;
    }
    return VL_OK;
}

// ---------------------------------------------------------------------------------------------- sampler test hooks
static int run_sampler(int which, uint64_t seed, uint64_t a, double b, uint64_t n, uint64_t *out) {
    uint64_t *dev = nullptr;
    if (cudaMalloc((void **)&dev, std::max<uint64_t>(n, 1) * 8) != cudaSuccess) return fail(nullptr, VL_ERR_CUDA, "no CUDA device / out of memory");
    unsigned g = (unsigned)std::min<uint64_t>((n + 255) / 256 + 1, 4096);
    if (which == 0) k_test_poisson<<<g, 256>>>(seed, b, n, dev);
    else if (which == 1) k_test_binomial<<<g, 256>>>(seed, a, b, n, dev, make_binv(a, b));
    else k_test_below<<<g, 256>>>(seed, a, n, dev);
    cudaError_t e = cudaMemcpy(out, dev, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(nullptr, VL_ERR_CUDA, "sampler kernel failed: %s", cudaGetErrorString(e));
    return VL_OK;
}
VL_API int vl_test_poisson_fast(uint64_t seed, double lambda, uint64_t n, uint64_t *out2n) {
    uint64_t *dev = nullptr;
    if (cudaMalloc((void **)&dev, std::max<uint64_t>(2 * n, 1) * 8) != cudaSuccess) return fail(nullptr, VL_ERR_CUDA, "no CUDA device / out of memory");
    k_test_poisson_fast<<<(unsigned)std::min<uint64_t>((n + 255) / 256 + 1, 4096), 256>>>(seed, lambda, n, dev);
    cudaError_t e = cudaMemcpy(out2n, dev, 2 * n * 8, cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(nullptr, VL_ERR_CUDA, "sampler kernel failed: %s", cudaGetErrorString(e));
    return VL_OK;
}
VL_API int vl_test_poisson(uint64_t seed, double lambda, uint64_t n, uint64_t *out) { return run_sampler(0, seed, 0, lambda, n, out); }
VL_API int vl_test_binomial(uint64_t seed, uint64_t trials, double p, uint64_t n, uint64_t *out) { return run_sampler(1, seed, trials, p, n, out); }
VL_API int vl_test_below(uint64_t seed, uint64_t bound, uint64_t n, uint64_t *out) { return run_sampler(2, seed, bound, 0.0, n, out); }
