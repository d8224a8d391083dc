/*
 * virolution_gpu.h — C ABI of the B200-native population-update path.
 *
 * Drop-in boundary for virolution's `trait Simulation<S>` (reference
 * src/simulation.rs:177-219).  A Rust `impl Simulation<Nt> for GpuSimulation`
 *  binds exactly these entry points; in this repository
 * the Python harness `virolution_b200/` (ctypes, `virolution_b200/abi.py`)
 * calls them.  Plain pointers and sizes only.
 *
 * Conventions
 *  - every function returns `int`: 0 = ok, <0 = error class (vl_status);
 *    `vl_last_error(sim)` returns the message of the last failure on that
 *    handle (reference: `VirolutionError`, src/errors.rs:5-15; the loop itself
 *    panics — those panics are VL_ERR_REFERENCE_PANIC here).
 *  - symbols (bases) are indices A=0,T=1,C=2,G=3 (src/encoding.rs:65-72).
 *  - fitness tables are row-major L x 4 f64 (src/core/fitness/table.rs:82-84).
 *  - transfer matrices are row-major, row = target, col = origin
 *    (src/config/schedule.rs:96-98).
 *  - host buffers unless the name says `_device`.
 *  - a handle owns one CUDA stream and binds its device on every call; calls
 *    on different handles may come from different threads (rayon workers in
 *    the reference, src/runner.rs:388-399).
 */
#ifndef VIROLUTION_GPU_H
#define VIROLUTION_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vl_sim vl_sim; /* one compartment: BasicSimulation, src/simulation.rs:221-232 */
typedef struct vl_pop vl_pop; /* detached Population<Store<S>>, src/core/population.rs:156-163 */

typedef enum vl_status {
    VL_OK = 0,
    VL_ERR_INITIALIZATION = -1,  /* VirolutionError::InitializationError */
    VL_ERR_IMPLEMENTATION = -2,  /* VirolutionError::ImplementationError */
    VL_ERR_CUDA = -3,            /* device/runtime failure (no reference analogue) */
    VL_ERR_CAPACITY = -4,        /* device arena / table capacity exhausted */
    VL_ERR_REFERENCE_PANIC = -5, /* the reference panics here (e.g. all-zero weights,
                                    src/core/population.rs:386) */
    VL_ERR_ARGUMENT = -6
} vl_status;

/* UtilityFunction, src/core/fitness/utility.rs:5-23 */
typedef enum vl_utility_kind { VL_UTILITY_LINEAR = 0, VL_UTILITY_ALGEBRAIC = 1, VL_UTILITY_HILL = 2 } vl_utility_kind;
typedef struct vl_utility {
    int32_t kind;
    int32_t _pad;
    double upper; /* Algebraic */
    double p, n;  /* Hill; k = (1-p)/p is derived (utility.rs:44) */
} vl_utility;

/* EpiEntry, src/core/fitness/epistasis.rs:15-21 (naturally aligned here; the
 * .npy record is packed — the harness converts). */
typedef struct vl_epi_entry {
    uint64_t pos1;
    uint64_t pos2;
    double value;
    uint8_t base1;
    uint8_t base2;
    uint8_t _pad[6];
} vl_epi_entry;

/* FitnessProvider, src/providers/fitness.rs:18-22,176-206 */
typedef enum vl_fitness_kind {
    VL_FITNESS_NONE = 0,          /* Option::None: host uses 1.0 (simulation.rs:62,141) */
    VL_FITNESS_NEUTRAL = 1,       /* Neutral: raw == 1 (providers/fitness.rs:140-147) */
    VL_FITNESS_NONEPISTATIC = 2,  /* table only */
    VL_FITNESS_EPISTATIC = 3      /* table x sparse pair table */
} vl_fitness_kind;
typedef struct vl_fitness {
    int32_t kind;
    int32_t _pad;
    const double *table;      /* n_sites*4, NULL for NONE/NEUTRAL */
    const vl_epi_entry *epi;  /* n_epi entries, EPISTATIC only */
    uint64_t n_epi;
    vl_utility utility;
} vl_fitness;

/* HostSpec{range, host}, src/core/hosts.rs:75-78; BasicHost providers simulation.rs:22-32 */
typedef struct vl_host_spec {
    uint64_t lo, hi; /* range lo..hi */
    vl_fitness reproductive;
    vl_fitness infective;
} vl_host_spec;

/* Parameters, src/config/parameters.rs:14-47 (host_model is resolved into specs) */
typedef struct vl_parameters {
    double mutation_rate;
    double recombination_rate;
    uint64_t host_population_size;
    double basic_reproductive_number;
    uint64_t max_population;
    double dilution;
    double substitution_matrix[16]; /* [from][to] */
    uint64_t n_hits;
} vl_parameters;

typedef struct vl_config {
    uint64_t n_sites;
    const uint8_t *wildtype; /* n_sites indices 0..3 */
    vl_parameters parameters;
    uint32_t n_specs;
    uint32_t compartment_id;       /* Philox stream id; distinct per compartment */
    const vl_host_spec *specs;
    uint64_t seed;                 /* the reference is unseeded (rand::rng()); this is ours */
    int32_t device;                /* CUDA ordinal */
    int32_t flags;                 /* VL_FLAG_* */
    uint64_t capacity_infectants;  /* 0 = auto (from max_population) */
    uint64_t capacity_haplotypes;  /* 0 = auto */
    uint64_t capacity_arena_words; /* 0 = auto; u32 words of packed mutation entries */
} vl_config;

enum {
    VL_FLAG_NO_BUCKETS = 4, /* build the host map with one histogram atomic per infectant instead of the
                               bucket-partitioned counting sort (same result; the bucketed path needs uniform host
                               draws and H >= 65536, this flag lets tests compare the two) */
    VL_FLAG_GENERAL_REPLICATE = 8, /* never use the fused replicate kernel (one pass over whole hosts per CTA); the
                                      general kernels give the same result, this flag lets tests compare the two */
    VL_FLAG_TREE_PLAN = 2, /* split the resampling draws over tiles with the conditional-binomial tree instead of
                              Poissonisation (both exact; the tree is the rare-case fallback, this flag lets tests
                              exercise it) */
    VL_FLAG_NO_TILES = 16 /* device-resident loops (vl_run, vl_next_generation) never take the tile-strided generation
                             (records appended per host tile, mapped fitness carried with the infectant); the general
                             kernels give the same populations, this flag lets tests compare the two */
};

/* ---- lifecycle -------------------------------------------------------- */
int vl_create(const vl_config *config, vl_sim **out);   /* BasicSimulation::new, simulation.rs:235-256 */
int vl_destroy(vl_sim *sim);
const char *vl_last_error(const vl_sim *sim);            /* NULL sim: last create error */
int vl_abi_version(void);

/* ---- Simulation trait (src/simulation.rs:177-219) ---------------------- */
int vl_increment_generation(vl_sim *sim);                /* :178 */
int vl_get_generation(const vl_sim *sim, uint64_t *out); /* :182 */
int vl_set_generation(vl_sim *sim, uint64_t generation); /* shared Arc<Generation>, runner.rs:72 */
int vl_set_parameters(vl_sim *sim, const vl_parameters *p); /* :355-363; like the reference it does
                                                             not reach the BasicHost copies */
int vl_population_len(vl_sim *sim, uint64_t *out);       /* get_population().len() */
int vl_set_population_wildtype(vl_sim *sim, uint64_t n); /* population![wt; n], runner.rs:263 */
int vl_infect(vl_sim *sim);                              /* :365-387 */
int vl_mutate_infectants(vl_sim *sim);                   /* :389-428 */
/* replicate_infectants, :430-485.  offspring_out may be NULL (kept on device). */
int vl_replicate_infectants(vl_sim *sim, uint64_t *offspring_out, uint64_t n);
/* offspring == NULL means "the device-resident map of the last replicate". */
int vl_target_size(vl_sim *sim, const uint64_t *offspring, uint64_t n, double *out);      /* :491-503 */
int vl_sample_indices(vl_sim *sim, const uint64_t *offspring, uint64_t n, uint64_t amount,
                      uint64_t *indices_out);                                            /* :505-514 */
int vl_subsample_population(vl_sim *sim, const uint64_t *offspring, uint64_t n, double factor,
                            vl_pop **out);                                               /* :516-527 */
int vl_select(vl_sim *sim, const uint64_t *indices, uint64_t n, vl_pop **out);            /* Population::select,
                                                                                           population.rs:406-430 */
/* set_population(Population::from_iter(parts)), :340-345 + population.rs:190-205. Consumes parts. */
int vl_set_population(vl_sim *sim, vl_pop *const *parts, uint32_t n_parts);
int vl_pop_len(const vl_pop *pop, uint64_t *out);
int vl_pop_free(vl_pop *pop);
int vl_next_generation(vl_sim *sim);                     /* default method, :200-218 */

/* ---- fused, device-resident loops (what Runner::run drives, runner.rs:336-423) */
/* n generations of a single compartment with the identity transfer; no host sync inside. */
int vl_run(vl_sim *sim, uint64_t n_generations);
/* increment_generation + infect + mutate_infectants + replicate_infectants as one enqueue, offspring map left on the
 * device: everything Runner::run does per compartment before the transmission block (runner.rs:382-399). */
int vl_advance_to_offspring(vl_sim *sim);
/* one generation over several compartments living in this process.
 * rule: 0 = parallel-build rule (runner.rs:401-416), 1 = serial-build rule (:550-631).
 * Failure: VL_ERR_CAPACITY from the size check means no compartment was changed (every new population is checked against
 * its capacity_infectants before any is installed).  Any other error may leave some compartments one generation ahead of
 * the others: the handles must then be discarded (vl_destroy), like a Runner whose loop panicked. */
int vl_step(vl_sim *const *sims, uint32_t n_sims, const double *transfer, int rule);
int vl_synchronize(vl_sim *sim);
/* sum over generations of population sizes at generation start since creation/reset. */
int vl_infectant_generations(vl_sim *sim, uint64_t *out, int reset);

/* ---- replay hooks: feed recorded draws instead of Philox --------------- */
/* infections[i] = host index or -1 (Option<usize>, hosts.rs:103). One-shot: consumed by the next vl_infect. */
int vl_replay_infections(vl_sim *sim, const int64_t *infections, uint64_t n);
/* per infectant CSR of (site, to-base) in the order the reference applies them (sorted sites);
 * consumed by the next vl_mutate_infectants. */
int vl_replay_mutations(vl_sim *sim, const uint64_t *offsets, uint64_t n, const uint32_t *sites,
                        const uint8_t *bases);
/* recombination events in processing order: host, slice positions i<j, window [s0,s1), which parent
 * slot is replaced (0 -> i, 1 -> j); consumed by the next vl_mutate_infectants. */
int vl_replay_recombinations(vl_sim *sim, uint64_t n_events, const uint64_t *host, const uint32_t *i,
                             const uint32_t *j, const uint32_t *s0, const uint32_t *s1,
                             const uint8_t *which);
/* offspring counts; consumed by the next vl_replicate_infectants (λ and host sums are still computed). */
int vl_replay_offspring(vl_sim *sim, const uint64_t *offspring, uint64_t n);

/* ---- inspection / export (parity tests, final.<i>.csv, sample writers) - */
int vl_get_infections(vl_sim *sim, int64_t *out, uint64_t n);          /* HostMapBuffer.infections */
int vl_get_host_map(vl_sim *sim, uint64_t *offsets_out /*H+1*/, uint64_t *hosts_out /*n infected*/,
                    uint64_t *n_infected_out);                          /* hosts.rs:166-192 */
/* mapped fitness f_i, host sum S and the Poisson mean lambda_i the last replicate used (0.0 where
 * Poisson::new fails, i.e. no offspring); NaN where not infected. */
int vl_get_replication_terms(vl_sim *sim, double *fitness_out, double *host_sum_out, double *lambda_out,
                             uint64_t n);
int vl_get_offspring_prefix(vl_sim *sim, uint64_t *prefix_out, uint64_t n); /* inclusive scan */
/* raw (pre-utility) fitness of every infectant under provider (spec, which): which 0=reproductive,1=infective */
int vl_get_raw_fitness(vl_sim *sim, uint32_t spec, int which, double *out, uint64_t n);
/* mutation sets: offsets_out[n+1]; then entries (pos, base) of get_mutations() sorted by pos.
 * Call with positions_out == NULL to get the total in *n_entries. */
int vl_export_mutations(vl_sim *sim, uint64_t *offsets_out, uint64_t n, uint32_t *positions_out,
                        uint8_t *bases_out, uint64_t *n_entries);
int vl_get_base(vl_sim *sim, uint64_t infectant, uint64_t position, uint8_t *out); /* Haplotype::get_base */
/* population statistics of the `stats` schedule event (src/stats/population.rs, src/runner.rs:634-673) */
int vl_stats_frequencies(vl_sim *sim, double *out /* n_sites x 4, row-major */, uint64_t n_sites);   /* :15-40 */
int vl_stats_mean_fitness(vl_sim *sim, uint32_t spec, int which, double *out);                       /* :50-56 */
int vl_get_haplotype_ids(vl_sim *sim, uint32_t *out, uint64_t n);      /* device-local ids (labels) */
/* counters: [0] live+dead haplotype records, [1] arena words used, [2] gc runs, [3] kernels launched */
int vl_get_counters(vl_sim *sim, uint64_t *out4);
/* capacities: [0] infectant slots, [1] haplotype records, [2] arena words, [3] generations between compaction checks
 * (a compaction runs when the record table or the arena is half full: the stand-in for Rc/Arc drops, src/references/sync.rs:107-115) */
int vl_get_capacities(vl_sim *sim, uint64_t *out4);
int vl_collect_garbage(vl_sim *sim);

/* ---- migrant exchange between handles / GPUs --------------------------- */
/* Pack a detached population into a self-contained byte image (haplotype payload + indices) in a
 * device buffer owned by the library; image_device/bytes stay valid until the pop is freed. */
int vl_pop_pack_device(vl_sim *origin, vl_pop *pop, void **image_device, uint64_t *bytes);
/* Build a detached population on `target` from a packed image located in target-device memory. */
int vl_pop_unpack_device(vl_sim *target, const void *image_device, uint64_t bytes, vl_pop **out);

/* ---- migrant exchange between processes (one process per GPU) ------------ */
/* The transmission block (src/runner.rs:401-422) when compartments live in different processes: the origin turns
 * its per-target parts into ONE flat image whose per-target ranges are contiguous, the caller moves the ranges
 * with its collective (NCCL all-to-all-v over NVLink in virolution_b200/distributed.py) and every target imports
 * what it received.  Pointers are device memory owned by the origin handle, valid until its next export. */
typedef struct vl_migrant_image {
    const uint32_t *len_device;     /* n_migrants: entries of each migrant's haplotype; 0x80000000 = wildtype */
    const double *raw_device;       /* n_migrants * n_providers memoised raw fitness, migrant-major */
    const uint32_t *entries_device; /* n_words flattened haplotype entries, migrant after migrant */
    uint64_t n_migrants, n_words, n_providers;
    uint64_t cap_migrants, cap_words; /* elements the buffers can hold before the library reallocates them (lets a caller
                                         cache its views of the buffers across generations) */
} vl_migrant_image;
int vl_migrants_export_device(vl_sim *origin, vl_pop *const *parts, uint32_t n_parts, vl_migrant_image *image,
                              uint64_t *migrants_per_part, uint64_t *words_per_part);
/* Buffers are device memory of the target's GPU, complete before the call (synchronise the stream that wrote them). */
int vl_migrants_import_device(vl_sim *target, const uint32_t *len_device, const double *raw_device,
                              const uint32_t *entries_device, uint64_t n_migrants, uint64_t n_words, vl_pop **out);
/* n subsamples of the device-resident offspring map, out[k] = subsample_population(&offspring, factors[k]);
 * one synchronisation for all of them (the per-target parts of one origin, src/runner.rs:404-416). */
int vl_subsample_populations(vl_sim *sim, const double *factors, uint32_t n, vl_pop **out);

/* ---- peer exchange: the transmission block fused with its transport --------------------------------------------
 * One compartment per process/GPU of one NVLink domain (src/runner.rs:401-416 with compartment c on rank c).  Every
 * rank owns a receive buffer that the other ranks map through CUDA IPC; per generation an origin writes its migrants
 * (length, raw fitness, entries) straight into the targets' buffers and raises a flag, the target waits on its flags
 * on the device and builds its new population.  Nothing is synchronised with or read back by the host.
 * `transfer` is row-major world x world, row = target, col = origin, fixed for the lifetime of the exchange; all ranks
 * must pass the same matrix, max_population, providers and words_per_migrant (the capacity of a block's entry area is
 * words_per_migrant x its migrant capacity; 0 = 64). */
typedef struct vl_exchange vl_exchange;
int vl_exchange_create(vl_sim *sim, uint32_t rank, uint32_t world, const double *transfer, uint64_t words_per_migrant,
                       int sharded, uint64_t shard_seed, vl_exchange **out,
                       void *ipc_handle_out /* 64 bytes: cudaIpcMemHandle_t of the receive buffer */);
/* sharded != 0: ONE compartment split over the ranks.  Rank g owns host_population_size hosts
 * (its handle is created with the LOCAL host count, the GLOBAL max_population and dilution, and a capacity_infectants
 * of about 1.3 x max_population / world) and the infectants that infect them.  Per generation every rank publishes its
 * offspring total to all ranks, every rank splits the N' = (min(W * dilution, max_population)) as usize draws of
 * Population::sample (src/simulation.rs:491-527) over the world^2 (target, origin) cells with the same multinomial
 * (cell weight W_origin: a new infectant's host, hence rank, is uniform — src/simulation.rs:365-387), and the cells
 * travel like migrants.  `transfer` is ignored (may be NULL); shard_seed must be the same on every rank. */
int vl_exchange_recv_buffer(vl_exchange *ex, void **ptr, uint64_t *bytes);
/* ipc_handles: world x 64 bytes gathered from all ranks (own slot ignored).  local_ptrs (optional): a non-NULL entry
 * is that rank's receive buffer in THIS process (vl_exchange_recv_buffer), used instead of its handle. */
int vl_exchange_connect(vl_exchange *ex, const void *ipc_handles, void *const *local_ptrs);
/* increment_generation .. replicate_infectants, then the transmission block, for this rank (runner.rs:382-422):
 * push = the generation up to the offspring map, the part of every target, the writes into the targets' buffers;
 * pull = device-side wait for every origin's block, import, set_population.  step = push + pull.  Ranks that share a
 * process push all before they pull any (a stream waiting on the device must not depend on work of the same process
 * that has not been submitted yet). */
int vl_exchange_begin(vl_exchange *ex); /* increment_generation .. replicate_infectants (+ totals of a sharded compartment) */
int vl_exchange_push(vl_exchange *ex);  /* calls vl_exchange_begin if the caller has not */
int vl_exchange_pull(vl_exchange *ex);
int vl_exchange_step(vl_exchange *ex);
/* n generations of vl_exchange_step enqueued back to back; every generation but the last also prepares the next one's
 * infection while it draws and imports (what vl_run does for a single device).  Same populations as n vl_exchange_step calls. */
int vl_exchange_run(vl_exchange *ex, uint64_t n_generations);
/* A new transfer matrix from the next vl_exchange_push on (the schedule's `transmission` events, src/config/schedule.rs:196-266).
 * Every entry must be <= the entry the exchange was created with: create it with the entry-wise maximum over the schedule. */
int vl_exchange_set_transfer(vl_exchange *ex, const double *transfer);
int vl_exchange_destroy(vl_exchange *ex);

/* ---- measurement hooks (bench.py): CUDA events on the handle's own stream ------------------ */
int vl_timer_start(vl_sim *sim);
int vl_timer_stop(vl_sim *sim, double *ms_out);            /* synchronises; device time since vl_timer_start */
int vl_profile(vl_sim *sim, int enable);                   /* bracket every kernel launch with events */
int vl_profile_report(vl_sim *sim, char *buf, uint64_t cap); /* "<kernel> <launches> <total_ms>\n" per kernel */
int vl_debug_last_generation(vl_sim *sim, uint64_t *offspring_out, int64_t *hosts_out, double *host_sums_out, double *fitness_out,
                             uint64_t n); /* diagnostic: what the last generation of vl_run / vl_next_generation left on the
                                             device, in the order of the population it STARTED with (n = its size): offspring
                                             counts, hosts, per-host fitness sums and carried mapped fitness.  Only after a
                                             device-resident generation without a host map (no reference analogue; tests) */
int vl_debug_stamps(vl_sim *sim, uint64_t *out8);          /* %globaltimer stamps of the last plan kernel's phases */

/* ---- device sampler hooks (distribution tests of the Philox-driven samplers that stand in for
 * rand_distr 0.5.1 Poisson / Binomial and rand 0.9.0 Uniform; call sites src/simulation.rs:47-48,148-151) */
int vl_test_poisson(uint64_t seed, double lambda, uint64_t n, uint64_t *out);
/* out2n[0..n) = the replicate kernels' f32-guarded Poisson inversion, out2n[n..2n) = the plain f64 inversion of the
 * same uniforms (lambda <= 0: a per-draw mean spread over (0,12)); the two halves must be equal. */
int vl_test_poisson_fast(uint64_t seed, double lambda, uint64_t n, uint64_t *out2n);
int vl_test_binomial(uint64_t seed, uint64_t trials, double p, uint64_t n, uint64_t *out);
int vl_test_below(uint64_t seed, uint64_t bound, uint64_t n, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* VIROLUTION_GPU_H */
