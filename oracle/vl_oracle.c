/*
 * vl_oracle.c — CPU restatement of virolution's per-generation population update.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (virolution_b200/, the C ABI in
 * include/virolution_gpu.h) may link, import or execute this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the
 * checker and as the reported CPU baseline.
 *
 * What it restates (all citations relative to /root/reference):
 *   src/simulation.rs:56-154,365-527      BasicHost::{infect,mutate,replicate}, BasicSimulation
 *   src/core/hosts.rs:166-192             HostMapBuffer::build (counting sort, back-fill)
 *   src/core/haplotype.rs:241-288,549-551,661-705,836-887   delta-encoded haplotype tree
 *   src/core/attributes.rs:196-241        raw value memo; utility applied on read
 *   src/providers/fitness.rs:212-228      get_fitness = raw(parent) * update | compute
 *   src/core/fitness/{table,epistasis,utility}.rs
 *   src/core/population.rs:385-430        Population::{sample,select}
 *   src/runner.rs:401-416 / :550-631      transmission, parallel-build and serial-build rules
 *
 * Third-party arithmetic (sources absent from /root/reference; Cargo.lock pins rand 0.9.0,
 * rand_distr 0.5.1) is restated from the crates' published algorithms:
 *   Poisson: Knuth product method for lambda < 12, Cauchy-envelope rejection otherwise;
 *   Binomial: BINV inversion for n*p < 10 (for n*p >= 10 this file sums BINV chunks instead of
 *             BTPE — same distribution, different stream);
 *   WeightedAliasIndex: Vose alias tables on integer weights;
 *   WeightedIndex: cumulative weights + partition point; Uniform/Bernoulli/choose.
 * The reference's generator is an OS-seeded ChaCha12 with no seed option, so bit-stream parity
 * with the Rust binary is impossible by construction; this file uses xoshiro256++ and records
 * every random decision (the "draw log") so the CUDA path can replay it bit-exactly.
 *
 * PARITY PIN STATUS: the reference's own known-answer tests for this path (table indexing,
 * CSR host map [8,6,2,0]/[9,7,4,1], descendant/recombinant sequences, merge-equivalence,
 * wildtype table entries) are reproduced in tests/test_oracle_kat.py.  No reference test pins
 * mutant fitness values, offspring counts or sample sizes, and the reference cannot be built
 * here (no Rust toolchain) — those outputs are "parity unpinned" beyond this restatement.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/virolution_gpu.h"

#define ORC_EXPORT __attribute__((visibility("default")))
#define MAX_PROVIDERS 64

/* ------------------------------------------------------------------ rng */
typedef struct { uint64_t s[4]; } Rng;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void rng_seed(Rng *r, uint64_t seed) {
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&seed);
}
static inline uint64_t rng_u64(Rng *r) {
    uint64_t *s = r->s;
    uint64_t result = rotl(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
/* rand: random::<f64>() = 53 random bits * 2^-53 */
static inline double rng_f64(Rng *r) { return (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
/* rand: Uniform<usize>::new(0,n) — widening multiply with rejection of the biased zone */
static inline uint64_t rng_below(Rng *r, uint64_t n) {
    uint64_t zone = (n << __builtin_clzll(n)) - 1;
    for (;;) {
        __uint128_t m = (__uint128_t)rng_u64(r) * n;
        if ((uint64_t)m <= zone) return (uint64_t)(m >> 64);
    }
}
/* rand: Bernoulli — p scaled to 2^64 */
static inline int rng_bernoulli(Rng *r, double p) {
    if (p >= 1.0) return 1;
    uint64_t p_int = (uint64_t)(p * 18446744073709551616.0);
    return rng_u64(r) < p_int;
}
/* rand_distr 0.5.1 Binomial (BINV branch; chunked for n*p >= 10, see header) */
static uint64_t binv(Rng *r, uint64_t n, double p) {
    const double q = 1.0 - p, s = p / q, a = (double)(n + 1) * s;
    for (;;) {
        double rr = pow(q, (double)n);
        double u = rng_f64(r);
        uint64_t x = 0;
        while (u > rr) {
            u -= rr;
            x += 1;
            if (x > 110) break;
            rr *= a / (double)x - s;
        }
        if (x <= 110) return x;
    }
}
static uint64_t rng_binomial(Rng *r, uint64_t n, double p) {
    if (p <= 0.0 || n == 0) return 0;
    if (p >= 1.0) return n;
    int flipped = p > 0.5;
    if (flipped) p = 1.0 - p;
    uint64_t x = 0;
    if ((double)n * p < 10.0) {
        x = binv(r, n, p);
    } else {
        uint64_t chunk = (uint64_t)(8.0 / p);
        if (chunk < 1) chunk = 1;
        uint64_t left = n;
        while (left > 0) {
            uint64_t c = left < chunk ? left : chunk;
            x += binv(r, c, p);
            left -= c;
        }
    }
    return flipped ? n - x : x;
}
/* rand_distr 0.5.1 Poisson::new + sample; returns 0 and sets *ok=0 where new() errs */
static double rng_poisson(Rng *r, double lambda, int *ok) {
    if (!isfinite(lambda) || !(lambda > 0.0) || lambda > 1.844e19) { *ok = 0; return 0.0; }
    *ok = 1;
    if (lambda < 12.0) {
        double exp_lambda = exp(-lambda);
        double result = 0.0, p = rng_f64(r);
        while (p > exp_lambda) { p *= rng_f64(r); result += 1.0; }
        return result;
    }
    const double log_lambda = log(lambda), sqrt_2lambda = sqrt(2.0 * lambda);
    const double magic_val = lambda * log_lambda - lgamma(1.0 + lambda);
    for (;;) {
        double result, comp_dev;
        for (;;) {
            comp_dev = tan(M_PI * rng_f64(r)); /* standard Cauchy */
            result = sqrt_2lambda * comp_dev + lambda;
            if (result >= 0.0) break;
        }
        result = floor(result);
        double check = 0.9 * (1.0 + comp_dev * comp_dev) *
                       exp(result * log_lambda - lgamma(1.0 + result) - magic_val);
        if (rng_f64(r) <= check) return result;
    }
}
/* rand WeightedIndex<f64>: cumulative weights (last dropped), chosen ~ U[0,total),
 * index = partition_point(cum <= chosen) */
static int rng_weighted4(Rng *r, const double *w) {
    double c0 = w[0], c1 = c0 + w[1], c2 = c1 + w[2], total = c2 + w[3];
    double chosen = rng_f64(r) * total;
    int idx = 0;
    if (c0 <= chosen) idx = 1;
    if (c1 <= chosen) idx = 2;
    if (c2 <= chosen) idx = 3;
    return idx;
}

/* Vose alias tables on integer weights (rand_distr WeightedAliasIndex<usize>) */
typedef struct { uint64_t n, weight_sum; uint64_t *odds; uint32_t *alias; } Alias;
static int alias_build(Alias *a, const uint64_t *w, uint64_t n) {
    a->n = n; a->odds = NULL; a->alias = NULL;
    if (n == 0) return -1;
    uint64_t sum = 0;
    for (uint64_t i = 0; i < n; i++) sum += w[i];
    if (sum == 0) return -1; /* Err(InsufficientNonZero) -> unwrap panics */
    a->weight_sum = sum;
    a->odds = (uint64_t *)malloc(n * sizeof(uint64_t));
    a->alias = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *small = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *big = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint64_t ns = 0, nb = 0;
    for (uint64_t i = 0; i < n; i++) {
        a->odds[i] = w[i] * n;
        a->alias[i] = (uint32_t)i;
        if (a->odds[i] < sum) small[ns++] = (uint32_t)i; else big[nb++] = (uint32_t)i;
    }
    while (ns > 0 && nb > 0) {
        uint32_t s = small[--ns], b = big[nb - 1];
        a->alias[s] = b;
        a->odds[b] -= sum - a->odds[s];
        if (a->odds[b] < sum) { nb--; small[ns++] = b; }
    }
    while (nb > 0) a->odds[big[--nb]] = sum;
    while (ns > 0) a->odds[small[--ns]] = sum;
    free(small); free(big);
    return 0;
}
static inline uint64_t alias_sample(const Alias *a, Rng *r) {
    uint64_t cand = rng_below(r, a->n);
    uint64_t u = rng_below(r, a->weight_sum);
    return u < a->odds[cand] ? cand : a->alias[cand];
}
static void alias_free(Alias *a) { free(a->odds); free(a->alias); }

/* ------------------------------------------------------------ haplotypes */
typedef struct { uint32_t pos; uint8_t from, to; } Change;
enum { K_WT = 0, K_MUT = 1, K_REC = 2 };
typedef struct {
    int64_t anc;    /* mutant ancestor | recombinant left ancestor */
    int64_t right;  /* recombinant right ancestor */
    uint32_t l, r;  /* recombinant window [l, r) */
    uint32_t n_ch;
    uint8_t kind;
    Change ch1;     /* SmallVec<[Change; 1]> inline slot (haplotype.rs:45,58) */
    Change *chs;    /* n_ch > 1 */
} Node;
#define CHUNK_BITS 16
#define CHUNK (1u << CHUNK_BITS)
#define MAX_CHUNKS (1u << 16)

typedef struct { uint32_t pos; uint8_t base; } Mut;
typedef struct { Mut *m; uint32_t n, cap; } MutSet;

typedef struct { uint64_t k1, k2; double v; uint64_t seq; } EpiDir;
typedef struct {
    int kind; /* vl_fitness_kind */
    double *table;
    EpiDir *epi; uint64_t n_epi;
    vl_utility util;
    int spec, which;
} Provider;
typedef struct { uint64_t lo, hi; int repro, infect; } Spec;

typedef struct { uint32_t infectant, site; uint8_t base; } MutLog;
typedef struct { uint64_t host; uint32_t spec, i, j, s0, s1; uint8_t which; } RecLog;

typedef struct {
    int64_t *pop; uint64_t n, cap;
    vl_parameters par;   /* BasicSimulation.parameters (updatable) */
    /* HostMapBuffer */
    int64_t *infections; uint64_t *offsets; uint64_t *hosts; uint64_t hcap, icap;
    /* logs of the last phase */
    MutLog *mlog; uint64_t n_mlog, cap_mlog;
    RecLog *rlog; uint64_t n_rlog, cap_rlog;
    double *fit, *hsum, *lam; /* per infectant terms of last replicate */
    uint64_t *offspring;
    uint64_t generation;
    Rng rng;
    uint64_t infectant_generations;
} Sim;

typedef struct orc_world {
    uint64_t L; uint8_t *wt;
    vl_parameters host_par; /* BasicHost.parameters (captured at construction) */
    Provider prov[MAX_PROVIDERS]; int n_prov;
    Spec *specs; uint32_t n_specs;
    Node *chunks[MAX_CHUNKS]; uint64_t n_nodes;
    double *raw[MAX_PROVIDERS][MAX_CHUNKS]; uint8_t *has[MAX_PROVIDERS][MAX_CHUNKS];
    Sim *sims; uint32_t n_sims;
    int n_threads;
    int mode; /* bit 0: no draw logs (timing runs: the reference keeps none); bit 1: "fully parallel CPU" variant — infect,
                 replicate and the sample draws run over all threads too (the reference's rayon build runs them serially) */
    Rng *trng; /* per-thread generators (rayon for_each_init(rand::rng)) */
    uint64_t **last_idx; uint64_t *last_idx_n; /* [t*C+o] sample indices of last step */
    char err[256];
} orc_world;

static inline Node *node_at(orc_world *w, int64_t id) { return &w->chunks[id >> CHUNK_BITS][id & (CHUNK - 1)]; }
static inline const Change *node_changes(const Node *n) { return n->n_ch > 1 ? n->chs : &n->ch1; }

static int64_t node_alloc(orc_world *w) {
    /* ids come from an atomic counter (the reference allocates nodes with the global allocator, no shared lock); the
     * tables grow a chunk at a time under a lock that is taken once per 2^CHUNK_BITS nodes */
    int64_t id = (int64_t)__atomic_fetch_add(&w->n_nodes, 1, __ATOMIC_RELAXED);
    uint64_t c = (uint64_t)id >> CHUNK_BITS;
    if (!__atomic_load_n(&w->chunks[c], __ATOMIC_ACQUIRE)) {
#pragma omp critical(orc_node_alloc)
        {
            if (!w->chunks[c]) {
                for (int p = 0; p < w->n_prov; p++) {
                    w->raw[p][c] = (double *)calloc(CHUNK, sizeof(double));
                    w->has[p][c] = (uint8_t *)calloc(CHUNK, 1);
                }
                Node *nc = (Node *)calloc(CHUNK, sizeof(Node));
                __atomic_store_n(&w->chunks[c], nc, __ATOMIC_RELEASE);
            }
        }
    }
    return id;
}

/* Haplotype::get_base (haplotype.rs:402-408,549-551,661-667,836-842) */
static uint8_t hap_get_base(orc_world *w, int64_t id, uint32_t pos) {
    for (;;) {
        Node *n = node_at(w, id);
        if (n->kind == K_WT) return w->wt[pos];
        if (n->kind == K_MUT) {
            const Change *c = node_changes(n);
            int found = -1;
            for (uint32_t k = 0; k < n->n_ch; k++) if (c[k].pos == pos) { found = (int)k; break; } /* first match */
            if (found >= 0) return c[found].to;
            id = n->anc;
        } else {
            id = (pos >= n->l && pos < n->r) ? n->anc : n->right;
        }
    }
}

/* Haplotype::create_descendant (haplotype.rs:241-260): from = parent.get_base(pos), assert from != to */
static int64_t hap_create_descendant(orc_world *w, int64_t parent, const uint32_t *pos, const uint8_t *to, uint32_t n, int *bad) {
    int64_t id = node_alloc(w);
    Node *nd = node_at(w, id);
    nd->kind = K_MUT; nd->anc = parent; nd->right = -1; nd->n_ch = n;
    Change *c = &nd->ch1;
    if (n > 1) { nd->chs = (Change *)malloc(n * sizeof(Change)); c = nd->chs; }
    for (uint32_t k = 0; k < n; k++) {
        c[k].pos = pos[k]; c[k].from = hap_get_base(w, parent, pos[k]); c[k].to = to[k];
        if (c[k].from == c[k].to && bad) *bad = 1; /* assert_ne!(from, to) */
    }
    return id;
}
/* Haplotype::create_recombinant (haplotype.rs:269-288) */
static int64_t hap_create_recombinant(orc_world *w, int64_t left, int64_t right, uint32_t l, uint32_t r) {
    int64_t id = node_alloc(w);
    Node *nd = node_at(w, id);
    nd->kind = K_REC; nd->anc = left; nd->right = right; nd->l = l; nd->r = r; nd->n_ch = 0;
    return id;
}

static void ms_reserve(MutSet *s, uint32_t n) {
    if (n > s->cap) { s->cap = n * 2 + 8; s->m = (Mut *)realloc(s->m, s->cap * sizeof(Mut)); }
}
static uint32_t ms_lower(const MutSet *s, uint32_t pos) {
    uint32_t lo = 0, hi = s->n;
    while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (s->m[mid].pos < pos) lo = mid + 1; else hi = mid; }
    return lo;
}
static int ms_get(const MutSet *s, uint32_t pos) {
    uint32_t i = ms_lower(s, pos);
    return (i < s->n && s->m[i].pos == pos) ? (int)s->m[i].base : -1;
}
static void ms_insert(MutSet *s, uint32_t pos, uint8_t base) {
    uint32_t i = ms_lower(s, pos);
    if (i < s->n && s->m[i].pos == pos) { s->m[i].base = base; return; }
    ms_reserve(s, s->n + 1);
    memmove(&s->m[i + 1], &s->m[i], (s->n - i) * sizeof(Mut));
    s->m[i].pos = pos; s->m[i].base = base; s->n++;
}
static void ms_remove(MutSet *s, uint32_t pos) {
    uint32_t i = ms_lower(s, pos);
    if (i < s->n && s->m[i].pos == pos) { memmove(&s->m[i], &s->m[i + 1], (s->n - i - 1) * sizeof(Mut)); s->n--; }
}
/* Haplotype::get_mutations (haplotype.rs:460-466,686-705,858-887); the HashMap becomes a
 * position-sorted set (iteration order only matters for f64 product order). */
static void hap_get_mutations(orc_world *w, int64_t id, MutSet *out) {
    out->n = 0;
    /* collect the mutant chain above id up to a wildtype or recombinant */
    int64_t stack_local[256]; int64_t *stack = stack_local; uint64_t sp = 0, scap = 256;
    int64_t cur = id;
    for (;;) {
        Node *n = node_at(w, cur);
        if (n->kind != K_MUT) break;
        if (sp == scap) {
            scap *= 2;
            int64_t *ns = (int64_t *)malloc(scap * sizeof(int64_t));
            memcpy(ns, stack, sp * sizeof(int64_t));
            if (stack != stack_local) free(stack);
            stack = ns;
        }
        stack[sp++] = cur;
        cur = n->anc;
    }
    Node *base = node_at(w, cur);
    if (base->kind == K_REC) {
        MutSet a = {0}, b = {0};
        hap_get_mutations(w, base->anc, &a);
        hap_get_mutations(w, base->right, &b);
        ms_reserve(out, a.n + b.n);
        for (uint32_t i = 0; i < a.n; i++) if (a.m[i].pos >= base->l && a.m[i].pos < base->r) ms_insert(out, a.m[i].pos, a.m[i].base);
        for (uint32_t i = 0; i < b.n; i++) if (b.m[i].pos < base->l || b.m[i].pos >= base->r) ms_insert(out, b.m[i].pos, b.m[i].base);
        free(a.m); free(b.m);
    }
    while (sp > 0) {
        Node *n = node_at(w, stack[--sp]);
        const Change *c = node_changes(n);
        for (uint32_t k = 0; k < n->n_ch; k++) {
            if (c[k].to == w->wt[c[k].pos]) ms_remove(out, c[k].pos); else ms_insert(out, c[k].pos, c[k].to);
        }
    }
    if (stack != stack_local) free(stack);
}

/* -------------------------------------------------------------- fitness */
static inline double tab(const Provider *p, uint32_t pos, uint8_t b) { return p->table[(uint64_t)pos * 4 + b]; }
/* UtilityFunction::apply (utility.rs:48-60) */
static double utility_apply(const vl_utility *u, double f) {
    switch (u->kind) {
    case VL_UTILITY_ALGEBRAIC: { double factor = 1. / (u->upper - 1.); return u->upper * factor * f / (1. + factor * f); }
    case VL_UTILITY_HILL: { double k = (1.0 - u->p) / u->p; double fp = pow(f, u->n); return fp / (k + fp); }
    default: return f;
    }
}
static int epi_cmp(const void *a, const void *b) {
    const EpiDir *x = (const EpiDir *)a, *y = (const EpiDir *)b;
    if (x->k1 != y->k1) return x->k1 < y->k1 ? -1 : 1;
    if (x->k2 != y->k2) return x->k2 < y->k2 ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
/* rows of the nested map: [lo,hi) range of directed entries with key k1 */
static void epi_row(const Provider *p, uint64_t k1, uint64_t *lo_out, uint64_t *hi_out) {
    uint64_t lo = 0, hi = p->n_epi;
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (p->epi[mid].k1 < k1) lo = mid + 1; else hi = mid; }
    uint64_t e = lo;
    while (e < p->n_epi && p->epi[e].k1 == k1) e++;
    *lo_out = lo; *hi_out = e;
}
static int epi_row_get(const Provider *p, uint64_t k1, uint64_t k2, double *v) {
    uint64_t lo, hi; epi_row(p, k1, &lo, &hi);
    if (lo == hi) return 0;
    uint64_t a = lo, b = hi;
    while (a < b) { uint64_t mid = (a + b) / 2; if (p->epi[mid].k2 < k2) a = mid + 1; else b = mid; }
    if (a < hi && p->epi[a].k2 == k2) { *v = p->epi[a].v; return 1; }
    return 0;
}
#define EKEY(pos, base) (((uint64_t)(pos) << 2) | (uint64_t)(base))

/* FitnessTable::update_fitness (table.rs:50-69) */
static double table_update(const Provider *p, const Node *n) {
    const Change *c = node_changes(n);
    double acc = 1.;
    for (uint32_t k = 0; k < n->n_ch; k++) acc = acc * tab(p, c[k].pos, c[k].to) / tab(p, c[k].pos, c[k].from);
    return acc;
}
/* EpistasisTable::update_fitness (epistasis.rs:105-173), quirks kept: the candidate filter tests
 * `to` twice and never `from`. */
static double epi_update(orc_world *w, const Provider *p, int64_t id) {
    Node *n = node_at(w, id);
    const Change *c = node_changes(n);
    int any = 0;
    for (uint32_t k = 0; k < n->n_ch; k++) { uint64_t lo, hi; epi_row(p, EKEY(c[k].pos, c[k].to), &lo, &hi); if (hi > lo) { any = 1; break; } }
    if (!any) return 1.;
    MutSet M = {0};
    hap_get_mutations(w, id, &M);
    double fitness = 1.;
    for (uint32_t k = 0; k < n->n_ch; k++) {
        uint64_t alo, ahi, rlo, rhi;
        epi_row(p, EKEY(c[k].pos, c[k].to), &alo, &ahi);
        if (ahi == alo) continue; /* not a candidate */
        epi_row(p, EKEY(c[k].pos, c[k].from), &rlo, &rhi);
        for (uint32_t m = 0; m < M.n; m++) {
            uint32_t pos = M.m[m].pos;
            if (pos <= c[k].pos) {
                int own = 0;
                for (uint32_t q = 0; q < n->n_ch; q++) if (c[q].pos == pos) { own = 1; break; }
                if (own) continue;
            }
            if (pos != c[k].pos) {
                double v;
                if (epi_row_get(p, EKEY(c[k].pos, c[k].to), EKEY(pos, M.m[m].base), &v)) fitness *= v;
                if (rhi > rlo && epi_row_get(p, EKEY(c[k].pos, c[k].from), EKEY(pos, M.m[m].base), &v)) fitness /= v;
            }
        }
    }
    free(M.m);
    return fitness;
}
/* FitnessTable::compute_fitness (table.rs:72-79) x EpistasisTable::compute_fitness
 * (epistasis.rs:176-193; the partner-absent rule is the reference's). */
static double full_compute(orc_world *w, const Provider *p, int64_t id) {
    MutSet M = {0};
    hap_get_mutations(w, id, &M);
    double acc = 1.;
    for (uint32_t m = 0; m < M.n; m++) acc = acc * tab(p, M.m[m].pos, M.m[m].base);
    if (p->kind == VL_FITNESS_EPISTATIC) {
        double fitness = 1.;
        for (uint32_t m = 0; m < M.n; m++) {
            uint64_t lo, hi; epi_row(p, EKEY(M.m[m].pos, M.m[m].base), &lo, &hi);
            for (uint64_t e = lo; e < hi; e++) {
                uint32_t q = (uint32_t)(p->epi[e].k2 >> 2); uint8_t cb = (uint8_t)(p->epi[e].k2 & 3);
                if (q > M.m[m].pos && ms_get(&M, q) != (int)cb) fitness *= p->epi[e].v;
            }
        }
        acc = acc * fitness;
    }
    free(M.m);
    return acc;
}
/* get_or_compute_raw + FitnessProvider::get_fitness (attributes.rs:196-217, providers/fitness.rs:212-228) */
static double raw_fitness(orc_world *w, int pi, int64_t id) {
    const Provider *p = &w->prov[pi];
    if (p->kind == VL_FITNESS_NEUTRAL) return 1.0;
    uint64_t c = (uint64_t)id >> CHUNK_BITS, o = (uint64_t)id & (CHUNK - 1);
    if (__atomic_load_n(&w->has[pi][c][o], __ATOMIC_ACQUIRE)) return w->raw[pi][c][o];
    Node *n = node_at(w, id);
    double v;
    if (n->kind == K_MUT) {
        double parent = raw_fitness(w, pi, n->anc);
        double update = table_update(p, n);
        if (p->kind == VL_FITNESS_EPISTATIC) update = update * epi_update(w, p, id);
        v = parent * update;
    } else {
        v = full_compute(w, p, id);
    }
    w->raw[pi][c][o] = v;
    __atomic_store_n(&w->has[pi][c][o], 1, __ATOMIC_RELEASE);
    return v;
}
static double mapped_fitness(orc_world *w, int pi, int64_t id) { return utility_apply(&w->prov[pi].util, raw_fitness(w, pi, id)); }

/* ------------------------------------------------------------ world API */
static int add_provider(orc_world *w, const vl_fitness *f, int spec, int which) {
    if (f->kind == VL_FITNESS_NONE) return -1;
    Provider *p = &w->prov[w->n_prov];
    memset(p, 0, sizeof(*p));
    p->kind = f->kind; p->util = f->utility; p->spec = spec; p->which = which;
    if (f->kind >= VL_FITNESS_NONEPISTATIC) {
        p->table = (double *)malloc(w->L * 4 * sizeof(double));
        memcpy(p->table, f->table, w->L * 4 * sizeof(double));
    }
    if (f->kind == VL_FITNESS_EPISTATIC) {
        /* EpistasisTable::from_vec (epistasis.rs:44-58): symmetric nested map, later inserts overwrite */
        EpiDir *d = (EpiDir *)malloc((2 * f->n_epi + 1) * sizeof(EpiDir));
        uint64_t nd = 0;
        for (uint64_t i = 0; i < f->n_epi; i++) {
            uint64_t a = EKEY(f->epi[i].pos1, f->epi[i].base1), b = EKEY(f->epi[i].pos2, f->epi[i].base2);
            d[nd].k1 = a; d[nd].k2 = b; d[nd].v = f->epi[i].value; d[nd].seq = nd; nd++;
            d[nd].k1 = b; d[nd].k2 = a; d[nd].v = f->epi[i].value; d[nd].seq = nd; nd++;
        }
        qsort(d, nd, sizeof(EpiDir), epi_cmp);
        uint64_t out = 0;
        for (uint64_t i = 0; i < nd; i++) {
            if (i + 1 < nd && d[i + 1].k1 == d[i].k1 && d[i + 1].k2 == d[i].k2) continue; /* keep last insert */
            d[out++] = d[i];
        }
        p->epi = d; p->n_epi = out;
    }
    return w->n_prov++;
}

ORC_EXPORT orc_world *orc_world_create(const vl_config *cfg, uint32_t n_compartments, int n_threads) {
    orc_world *w = (orc_world *)calloc(1, sizeof(orc_world));
    w->L = cfg->n_sites;
    w->wt = (uint8_t *)malloc(w->L);
    memcpy(w->wt, cfg->wildtype, w->L);
    w->host_par = cfg->parameters;
    w->n_specs = cfg->n_specs;
    w->specs = (Spec *)calloc(cfg->n_specs, sizeof(Spec));
    for (uint32_t s = 0; s < cfg->n_specs; s++) {
        w->specs[s].lo = cfg->specs[s].lo; w->specs[s].hi = cfg->specs[s].hi;
        w->specs[s].repro = add_provider(w, &cfg->specs[s].reproductive, (int)s, 0);
        w->specs[s].infect = add_provider(w, &cfg->specs[s].infective, (int)s, 1);
    }
    int maxt = 1;
#ifdef _OPENMP
    maxt = omp_get_max_threads();
#endif
    w->n_threads = n_threads > 0 ? n_threads : maxt;
    w->trng = (Rng *)malloc((size_t)(w->n_threads + 1) * sizeof(Rng));
    for (int t = 0; t <= w->n_threads; t++) rng_seed(&w->trng[t], cfg->seed * 0x9E3779B97F4A7C15ULL + 1000003ULL * (uint64_t)(t + 1));
    w->n_sims = n_compartments;
    w->sims = (Sim *)calloc(n_compartments, sizeof(Sim));
    for (uint32_t c = 0; c < n_compartments; c++) {
        w->sims[c].par = cfg->parameters;
        rng_seed(&w->sims[c].rng, cfg->seed ^ (0xA5A5A5A5ULL * (c + 1)));
    }
    w->last_idx = (uint64_t **)calloc((size_t)n_compartments * n_compartments, sizeof(uint64_t *));
    w->last_idx_n = (uint64_t *)calloc((size_t)n_compartments * n_compartments, sizeof(uint64_t));
    int64_t wt = node_alloc(w); /* node 0 = wildtype */
    node_at(w, wt)->kind = K_WT;
    return w;
}
ORC_EXPORT void orc_world_destroy(orc_world *w) {
    if (!w) return;
    for (uint64_t c = 0; c < MAX_CHUNKS && w->chunks[c]; c++) {
        uint64_t lim = CHUNK;
        for (uint64_t i = 0; i < lim; i++) if (w->chunks[c][i].n_ch > 1) free(w->chunks[c][i].chs);
        free(w->chunks[c]);
        for (int p = 0; p < w->n_prov; p++) { free(w->raw[p][c]); free(w->has[p][c]); }
    }
    for (int p = 0; p < w->n_prov; p++) { free(w->prov[p].table); free(w->prov[p].epi); }
    for (uint32_t c = 0; c < w->n_sims; c++) {
        Sim *s = &w->sims[c];
        free(s->pop); free(s->infections); free(s->offsets); free(s->hosts); free(s->mlog); free(s->rlog);
        free(s->fit); free(s->hsum); free(s->lam); free(s->offspring);
    }
    for (uint32_t i = 0; i < w->n_sims * w->n_sims; i++) free(w->last_idx[i]);
    free(w->last_idx); free(w->last_idx_n);
    free(w->sims); free(w->specs); free(w->wt); free(w->trng); free(w);
}
ORC_EXPORT const char *orc_last_error(orc_world *w) { return w->err; }

static void sim_reserve(Sim *s, uint64_t n, uint64_t H) {
    if (n > s->cap) { s->cap = n + n / 4 + 16; s->pop = (int64_t *)realloc(s->pop, s->cap * sizeof(int64_t)); }
    if (n > s->icap) {
        s->icap = n + n / 4 + 16;
        s->infections = (int64_t *)realloc(s->infections, s->icap * sizeof(int64_t));
        s->hosts = (uint64_t *)realloc(s->hosts, s->icap * sizeof(uint64_t));
        s->fit = (double *)realloc(s->fit, s->icap * sizeof(double));
        s->hsum = (double *)realloc(s->hsum, s->icap * sizeof(double));
        s->lam = (double *)realloc(s->lam, s->icap * sizeof(double));
        s->offspring = (uint64_t *)realloc(s->offspring, s->icap * sizeof(uint64_t));
    }
    if (H + 1 > s->hcap) { s->hcap = H + 1; s->offsets = (uint64_t *)realloc(s->offsets, s->hcap * sizeof(uint64_t)); }
}
ORC_EXPORT int orc_set_population_wildtype(orc_world *w, uint32_t c, uint64_t n) {
    Sim *s = &w->sims[c];
    sim_reserve(s, n, s->par.host_population_size);
    for (uint64_t i = 0; i < n; i++) s->pop[i] = 0;
    s->n = n;
    return 0;
}
ORC_EXPORT uint64_t orc_population_len(orc_world *w, uint32_t c) { return w->sims[c].n; }
ORC_EXPORT void orc_set_parameters(orc_world *w, uint32_t c, const vl_parameters *p) { w->sims[c].par = *p; } /* simulation.rs:355-363 */
ORC_EXPORT void orc_increment_generation(orc_world *w, uint32_t c) { w->sims[c].generation++; }
ORC_EXPORT uint64_t orc_get_generation(orc_world *w, uint32_t c) { return w->sims[c].generation; }
ORC_EXPORT uint64_t orc_infectant_generations(orc_world *w, uint32_t c) { return w->sims[c].infectant_generations; }

static const Spec *spec_of_host(orc_world *w, uint64_t h, uint32_t *idx) { /* hosts.rs:62-64 */
    for (uint32_t s = 0; s < w->n_specs; s++) if (h >= w->specs[s].lo && h < w->specs[s].hi) { if (idx) *idx = s; return &w->specs[s]; }
    return NULL;
}

/* HostMapBuffer::build counting sort (hosts.rs:166-192) over s->infections */
static void build_host_map(Sim *s, uint64_t H) {
    for (uint64_t h = 0; h <= H; h++) s->offsets[h] = 0;
    for (uint64_t i = 0; i < s->n; i++) if (s->infections[i] >= 0) s->offsets[s->infections[i]] += 1;
    for (uint64_t h = 0; h < H; h++) s->offsets[h + 1] += s->offsets[h];
    for (uint64_t i = 0; i < s->n; i++) if (s->infections[i] >= 0) { uint64_t h = (uint64_t)s->infections[i]; s->hosts[s->offsets[h] - 1] = i; s->offsets[h] -= 1; }
}

/* BasicSimulation::infect (simulation.rs:365-387) + BasicHost::infect (:56-65) */
ORC_EXPORT int orc_infect(orc_world *w, uint32_t c) {
    Sim *s = &w->sims[c];
    uint64_t H = s->par.host_population_size;
    sim_reserve(s, s->n, H);
    s->infectant_generations += s->n;
    const int par_all = (w->mode & 2) && w->n_threads > 1 && !omp_in_parallel();
#pragma omp parallel for schedule(static) num_threads(w->n_threads) if (par_all)
    for (uint64_t i = 0; i < s->n; i++) {
        Rng *rng = par_all ? &w->trng[omp_get_thread_num()] : &s->rng;
        int64_t inf = -1;
        for (uint64_t hit = 0; hit < s->par.n_hits; hit++) {
            uint64_t cand = rng_below(rng, H);
            const Spec *sp = spec_of_host(w, cand, NULL);
            if (!sp) continue;
            double infectivity = sp->infect >= 0 ? mapped_fitness(w, sp->infect, s->pop[i]) : 1.;
            if (rng_f64(rng) < infectivity) { inf = (int64_t)cand; break; }
        }
        s->infections[i] = inf;
    }
    build_host_map(s, H);
    return 0;
}
/* test hook: impose infections (replaying a log into the oracle itself / KAT of hosts.rs:294-331) */
ORC_EXPORT int orc_set_infections(orc_world *w, uint32_t c, const int64_t *inf, uint64_t n) {
    Sim *s = &w->sims[c];
    if (n != s->n) { s->n = n; }
    sim_reserve(s, n, s->par.host_population_size);
    memcpy(s->infections, inf, n * sizeof(int64_t));
    build_host_map(s, s->par.host_population_size);
    return 0;
}

static void push_mlog(Sim *s, uint32_t infectant, uint32_t site, uint8_t base) {
    if (s->n_mlog == s->cap_mlog) { s->cap_mlog = s->cap_mlog * 2 + 1024; s->mlog = (MutLog *)realloc(s->mlog, s->cap_mlog * sizeof(MutLog)); }
    s->mlog[s->n_mlog].infectant = infectant; s->mlog[s->n_mlog].site = site; s->mlog[s->n_mlog].base = base; s->n_mlog++;
}
static int cmp_u32(const void *a, const void *b) { uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return x < y ? -1 : x > y; }

/* BasicHost::mutate (simulation.rs:67-118) on one host slice */
static void host_mutate(orc_world *w, Sim *s, uint32_t spec_idx, uint64_t host, const uint64_t *ids, uint64_t k, Rng *rng) {
    const vl_parameters *hp = &w->host_par;
    int64_t local_buf[64];
    int64_t *g = k <= 64 ? local_buf : (int64_t *)malloc(k * sizeof(int64_t));
    for (uint64_t a = 0; a < k; a++) g[a] = s->pop[ids[a]];
    if (k > 1 && hp->recombination_rate > 0.) {
        for (uint64_t i = 0; i < k; i++) for (uint64_t j = i + 1; j < k; j++) { /* tuple_combinations */
            if (!rng_bernoulli(rng, hp->recombination_rate)) continue;
            uint32_t s0 = (uint32_t)rng_below(rng, w->L), s1 = (uint32_t)rng_below(rng, w->L);
            while (s0 == s1) s1 = (uint32_t)rng_below(rng, w->L);
            if (s0 > s1) { uint32_t t = s0; s0 = s1; s1 = t; }
            int64_t rec = hap_create_recombinant(w, g[i], g[j], s0, s1);
            uint8_t which = (uint8_t)rng_below(rng, 2); /* [i, j].choose(rng) */
            g[which ? j : i] = rec;
            if (!(w->mode & 1))
#pragma omp critical(orc_log)
            {
                if (s->n_rlog == s->cap_rlog) { s->cap_rlog = s->cap_rlog * 2 + 64; s->rlog = (RecLog *)realloc(s->rlog, s->cap_rlog * sizeof(RecLog)); }
                RecLog *r = &s->rlog[s->n_rlog++];
                r->host = host; r->spec = spec_idx; r->i = (uint32_t)i; r->j = (uint32_t)j; r->s0 = s0; r->s1 = s1; r->which = which;
            }
        }
    }
    for (uint64_t a = 0; a < k; a++) {
        uint64_t nm = rng_binomial(rng, w->L, hp->mutation_rate);
        if (nm == 0) continue;
        uint32_t sl[16]; uint8_t bl[16];
        uint32_t *sites = nm <= 16 ? sl : (uint32_t *)malloc(nm * sizeof(uint32_t));
        uint8_t *bases = nm <= 16 ? bl : (uint8_t *)malloc(nm);
        for (uint64_t m = 0; m < nm; m++) sites[m] = (uint32_t)rng_below(rng, w->L); /* with replacement */
        qsort(sites, nm, sizeof(uint32_t), cmp_u32);
        for (uint64_t m = 0; m < nm; m++) {
            uint8_t cur = hap_get_base(w, g[a], sites[m]); /* on the pre-mutation haplotype */
            bases[m] = (uint8_t)rng_weighted4(rng, &hp->substitution_matrix[cur * 4]);
        }
        g[a] = hap_create_descendant(w, g[a], sites, bases, (uint32_t)nm, NULL);
        if (!(w->mode & 1))
#pragma omp critical(orc_log)
        { for (uint64_t m = 0; m < nm; m++) push_mlog(s, (uint32_t)ids[a], sites[m], bases[m]); }
        if (nm > 16) { free(sites); free(bases); }
    }
    for (uint64_t a = 0; a < k; a++) s->pop[ids[a]] = g[a];
    if (g != local_buf) free(g);
}

/* BasicSimulation::mutate_infectants (simulation.rs:389-428); parallel over hosts like the
 * rayon build (per-worker generator), serial otherwise */
ORC_EXPORT int orc_mutate(orc_world *w, uint32_t c) {
    Sim *s = &w->sims[c];
    s->n_mlog = 0; s->n_rlog = 0;
    for (uint32_t si = 0; si < w->n_specs; si++) {
        const Spec *sp = &w->specs[si];
        uint64_t lo = sp->lo, hi = sp->hi;
        if (hi > s->par.host_population_size) hi = s->par.host_population_size; /* get_slice bound */
        if (w->n_threads > 1) {
#ifdef _OPENMP
            int nested = omp_in_parallel();
#else
            int nested = 1;
#endif
            if (!nested) {
#pragma omp parallel for schedule(dynamic, 4096) num_threads(w->n_threads)
                for (uint64_t h = lo; h < hi; h++) {
#ifdef _OPENMP
                    Rng *rng = &w->trng[omp_get_thread_num()];
#else
                    Rng *rng = &w->trng[0];
#endif
                    uint64_t b = s->offsets[h], e = s->offsets[h + 1];
                    if (e > b) host_mutate(w, s, si, h, &s->hosts[b], e - b, rng);
                }
                continue;
            }
        }
        for (uint64_t h = lo; h < hi; h++) {
            uint64_t b = s->offsets[h], e = s->offsets[h + 1];
            if (e > b) host_mutate(w, s, si, h, &s->hosts[b], e - b, &s->rng);
        }
    }
    return 0;
}

/* BasicSimulation::replicate_infectants (:430-485) + BasicHost::replicate (:120-154) */
ORC_EXPORT int orc_replicate(orc_world *w, uint32_t c, uint64_t *offspring_out) {
    Sim *s = &w->sims[c];
    const double R0 = w->host_par.basic_reproductive_number;
    for (uint64_t i = 0; i < s->n; i++) { s->offspring[i] = 0; s->fit[i] = NAN; s->hsum[i] = NAN; s->lam[i] = NAN; }
    for (uint32_t si = 0; si < w->n_specs; si++) {
        const Spec *sp = &w->specs[si];
        uint64_t hi = sp->hi < s->par.host_population_size ? sp->hi : s->par.host_population_size;
        const int par_all = (w->mode & 2) && w->n_threads > 1 && !omp_in_parallel();
#pragma omp parallel for schedule(static) num_threads(w->n_threads) if (par_all)
        for (uint64_t h = sp->lo; h < hi; h++) {
            Rng *prng = par_all ? &w->trng[omp_get_thread_num()] : &s->rng;
            uint64_t b = s->offsets[h], e = s->offsets[h + 1];
            if (e == b) continue;
            double sum = 0.0;
            for (uint64_t a = b; a < e; a++) {
                uint64_t i = s->hosts[a];
                double f = sp->repro >= 0 ? mapped_fitness(w, sp->repro, s->pop[i]) : 1.;
                s->fit[i] = f;
                sum += f; /* iter().sum() in slice order */
            }
            for (uint64_t a = b; a < e; a++) {
                uint64_t i = s->hosts[a];
                double lam = s->fit[i] * R0 / sum;
                s->hsum[i] = sum; s->lam[i] = lam;
                int ok; double x = rng_poisson(prng, lam, &ok);
                if (ok) s->offspring[i] = (uint64_t)x;
                else { uint64_t bits; memcpy(&bits, &s->fit[i], 8); s->offspring[i] = bits; } /* float view left in place */
            }
        }
    }
    if (offspring_out) memcpy(offspring_out, s->offspring, s->n * sizeof(uint64_t));
    return 0;
}

/* target_size (:491-503) */
ORC_EXPORT double orc_target_size(orc_world *w, uint32_t c, const uint64_t *off, uint64_t n) {
    Sim *s = &w->sims[c];
    if (n == 0) return 0.0;
    uint64_t sum = 0;
    for (uint64_t i = 0; i < n; i++) sum += off[i];
    double a = (double)sum * s->par.dilution, b = (double)s->par.max_population;
    return a <= b ? a : b; /* min_by partial_cmp: returns a on Equal/Less */
}
static uint64_t f64_as_usize(double x) { if (!(x > 0)) return 0; if (x >= 18446744073709551615.0) return UINT64_MAX; return (uint64_t)x; }

/* sample_indices (:505-514): returns -1 where the reference panics (all-zero weights) */
ORC_EXPORT int orc_sample_indices(orc_world *w, uint32_t c, const uint64_t *off, uint64_t n, uint64_t amount, uint64_t *out) {
    Sim *s = &w->sims[c];
    if (n == 0) return 0;
    Alias a;
    if (alias_build(&a, off, n) != 0) { snprintf(w->err, sizeof w->err, "WeightedAliasIndex::new: all weights zero"); return -1; }
    const int par_all = (w->mode & 2) && w->n_threads > 1 && !omp_in_parallel();
#pragma omp parallel for schedule(static) num_threads(w->n_threads) if (par_all)
    for (uint64_t k = 0; k < amount; k++) out[k] = alias_sample(&a, par_all ? &w->trng[omp_get_thread_num()] : &s->rng);
    alias_free(&a);
    return 0;
}
/* subsample size of subsample_population (:516-527) */
ORC_EXPORT uint64_t orc_subsample_size(orc_world *w, uint32_t c, const uint64_t *off, uint64_t n, double factor) {
    if (n == 0) return 0;
    return f64_as_usize(factor * orc_target_size(w, c, off, n));
}

/* detached populations are plain id vectors */
typedef struct { int64_t *ids; uint64_t n; } OPop;
ORC_EXPORT OPop *orc_select(orc_world *w, uint32_t c, const uint64_t *idx, uint64_t n) { /* population.rs:406-430 */
    Sim *s = &w->sims[c];
    OPop *p = (OPop *)malloc(sizeof(OPop));
    p->n = n; p->ids = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
    for (uint64_t k = 0; k < n; k++) p->ids[k] = s->pop[idx[k]];
    return p;
}
/* subsample_population = sample_indices + select (population.rs:385-403); NULL where the reference panics.
 * idx_out (optional, capacity = subsample size) receives the drawn indices. */
ORC_EXPORT OPop *orc_subsample(orc_world *w, uint32_t c, const uint64_t *off, uint64_t n, double factor, uint64_t *idx_out) {
    if (n == 0) { OPop *p = (OPop *)malloc(sizeof(OPop)); p->n = 0; p->ids = (int64_t *)malloc(8); return p; }
    uint64_t size = orc_subsample_size(w, c, off, n, factor);
    uint64_t *idx = idx_out ? idx_out : (uint64_t *)malloc((size ? size : 1) * sizeof(uint64_t));
    OPop *p = NULL;
    if (orc_sample_indices(w, c, off, n, size, idx) == 0) p = orc_select(w, c, idx, size);
    if (!idx_out) free(idx);
    return p;
}
ORC_EXPORT uint64_t orc_pop_len(OPop *p) { return p->n; }
ORC_EXPORT void orc_pop_free(OPop *p) { if (p) { free(p->ids); free(p); } }
/* set_population(Population::from_iter(parts)) — consumes parts */
ORC_EXPORT int orc_set_population(orc_world *w, uint32_t c, OPop **parts, uint32_t n_parts) {
    Sim *s = &w->sims[c];
    uint64_t total = 0;
    for (uint32_t k = 0; k < n_parts; k++) total += parts[k]->n;
    int64_t *np = (int64_t *)malloc((total ? total : 1) * sizeof(int64_t));
    uint64_t o = 0;
    for (uint32_t k = 0; k < n_parts; k++) { memcpy(np + o, parts[k]->ids, parts[k]->n * sizeof(int64_t)); o += parts[k]->n; orc_pop_free(parts[k]); }
    free(s->pop); s->pop = np; s->n = total; s->cap = total ? total : 1;
    sim_reserve(s, total, s->par.host_population_size);
    return 0;
}
/* Simulation::next_generation (:200-218) */
ORC_EXPORT int orc_next_generation(orc_world *w, uint32_t c) {
    Sim *s = &w->sims[c];
    s->generation++;
    if (s->n == 0) return 0;
    orc_infect(w, c); orc_mutate(w, c); orc_replicate(w, c, NULL);
    OPop *p = orc_subsample(w, c, s->offspring, s->n, 1., NULL);
    if (!p) return -1;
    return orc_set_population(w, c, &p, 1);
}

/* One generation over all compartments as Runner::run drives it.
 * rule 0: parallel build (runner.rs:382-422); rule 1: serial build (:536-631).
 * Sample indices per (target, origin) are kept for the draw log. */
ORC_EXPORT int orc_step(orc_world *w, const double *T, int rule) {
    const uint32_t C = w->n_sims;
    int fail = 0;
    for (uint32_t c = 0; c < C; c++) w->sims[c].generation++;
#pragma omp parallel for schedule(dynamic, 1) num_threads(w->n_threads) if (C > 1 && w->n_threads > 1)
    for (uint32_t c = 0; c < C; c++) orc_infect(w, c);
#pragma omp parallel for schedule(dynamic, 1) num_threads(w->n_threads) if (C > 1 && w->n_threads > 1)
    for (uint32_t c = 0; c < C; c++) { orc_mutate(w, c); orc_replicate(w, c, NULL); }
    OPop **parts = (OPop **)calloc((size_t)C * C, sizeof(OPop *));
    if (rule == 0) {
        uint64_t *pair_seed = (uint64_t *)malloc((size_t)C * C * sizeof(uint64_t));
        for (uint32_t k = 0; k < C * C; k++) pair_seed[k] = rng_u64(&w->trng[w->n_threads]);
#pragma omp parallel for schedule(dynamic, 1) num_threads(w->n_threads) if (C > 1 && w->n_threads > 1)
        for (uint32_t t = 0; t < C; t++) {
            for (uint32_t o = 0; o < C; o++) {
                Sim *so = &w->sims[o];
                uint64_t size = orc_subsample_size(w, o, so->offspring, so->n, T[t * C + o]);
                uint64_t *idx = (uint64_t *)malloc((size ? size : 1) * sizeof(uint64_t));
                /* the origin's generator is shared between target threads in this restatement; give
                 * each (t,o) pair its own stream to stay race-free */
                Sim tmp = *so; rng_seed(&tmp.rng, pair_seed[t * C + o]);
                OPop *p;
                if (so->n == 0) { p = (OPop *)malloc(sizeof(OPop)); p->n = 0; p->ids = (int64_t *)malloc(8); size = 0; }
                else {
                    Alias a;
                    if (alias_build(&a, so->offspring, so->n) != 0) { fail = 1; p = (OPop *)malloc(sizeof(OPop)); p->n = 0; p->ids = (int64_t *)malloc(8); size = 0; }
                    else { for (uint64_t k = 0; k < size; k++) idx[k] = alias_sample(&a, &tmp.rng); alias_free(&a); p = orc_select(w, o, idx, size); }
                }
                free(w->last_idx[t * C + o]); w->last_idx[t * C + o] = idx; w->last_idx_n[t * C + o] = size;
                parts[t * C + o] = p;
            }
        }
        free(pair_seed);
    } else {
        double *ts = (double *)malloc(C * sizeof(double));
        uint64_t *amount = (uint64_t *)calloc((size_t)C * C, sizeof(uint64_t));
        for (uint32_t t = 0; t < C; t++) ts[t] = orc_target_size(w, t, w->sims[t].offspring, w->sims[t].n);
        for (uint32_t t = 0; t < C; t++) for (uint32_t o = 0; o < C; o++) {
            if (t == o) continue;
            double m = ts[t] * T[t * C + o];
            uint64_t residue = rng_f64(&w->sims[t].rng) < m - floor(m) ? 1 : 0;
            amount[t * C + o] = f64_as_usize(m) + residue;
        }
        for (uint32_t t = 0; t < C && !fail; t++) {
            uint64_t mig = 0;
            for (uint32_t o = 0; o < C; o++) mig += amount[t * C + o];
            if (f64_as_usize(ts[t]) < mig) { fail = 1; break; } /* usize underflow panics */
            amount[t * C + t] = f64_as_usize(ts[t]) - mig;
        }
        for (uint32_t t = 0; t < C && !fail; t++) for (uint32_t o = 0; o < C; o++) {
            Sim *so = &w->sims[o];
            uint64_t size = so->n == 0 ? 0 : amount[t * C + o];
            uint64_t *idx = (uint64_t *)malloc((size ? size : 1) * sizeof(uint64_t));
            if (so->n && orc_sample_indices(w, o, so->offspring, so->n, size, idx) != 0) { fail = 1; size = 0; }
            free(w->last_idx[t * C + o]); w->last_idx[t * C + o] = idx; w->last_idx_n[t * C + o] = size;
            parts[t * C + o] = orc_select(w, o, idx, size);
        }
        free(ts); free(amount);
    }
    if (fail) {
        for (uint32_t k = 0; k < C * C; k++) orc_pop_free(parts[k]);
        free(parts);
        snprintf(w->err, sizeof w->err, "reference panics in transmission (extinct origin or migrant underflow)");
        return -1;
    }
    for (uint32_t t = 0; t < C; t++) orc_set_population(w, t, &parts[t * C], C);
    free(parts);
    return 0;
}
ORC_EXPORT uint64_t orc_last_indices(orc_world *w, uint32_t t, uint32_t o, uint64_t *out) {
    uint64_t n = w->last_idx_n[t * w->n_sims + o];
    if (out) memcpy(out, w->last_idx[t * w->n_sims + o], n * sizeof(uint64_t));
    return n;
}

/* ------------------------------------------------------- logs & export */
ORC_EXPORT void orc_get_infections(orc_world *w, uint32_t c, int64_t *out) { memcpy(out, w->sims[c].infections, w->sims[c].n * sizeof(int64_t)); }
ORC_EXPORT uint64_t orc_get_host_map(orc_world *w, uint32_t c, uint64_t *offsets, uint64_t *hosts) {
    Sim *s = &w->sims[c];
    uint64_t H = s->par.host_population_size;
    memcpy(offsets, s->offsets, (H + 1) * sizeof(uint64_t));
    memcpy(hosts, s->hosts, s->offsets[H] * sizeof(uint64_t));
    return s->offsets[H];
}
/* mutation log as CSR by infectant (entries keep the applied = sorted-site order); sites NULL -> count */
ORC_EXPORT uint64_t orc_get_mutation_log(orc_world *w, uint32_t c, uint64_t *offsets, uint32_t *sites, uint8_t *bases) {
    Sim *s = &w->sims[c];
    if (!sites) return s->n_mlog;
    /* stable by infectant: entries of one infectant were pushed contiguously in order */
    MutLog *tmp = (MutLog *)malloc((s->n_mlog ? s->n_mlog : 1) * sizeof(MutLog));
    uint64_t *cnt = (uint64_t *)calloc(s->n + 1, sizeof(uint64_t));
    for (uint64_t k = 0; k < s->n_mlog; k++) cnt[s->mlog[k].infectant + 1]++;
    for (uint64_t i = 0; i < s->n; i++) cnt[i + 1] += cnt[i];
    memcpy(offsets, cnt, (s->n + 1) * sizeof(uint64_t));
    for (uint64_t k = 0; k < s->n_mlog; k++) tmp[cnt[s->mlog[k].infectant]++] = s->mlog[k];
    for (uint64_t k = 0; k < s->n_mlog; k++) { sites[k] = tmp[k].site; bases[k] = tmp[k].base; }
    free(tmp); free(cnt);
    return s->n_mlog;
}
static int cmp_rlog(const void *a, const void *b) {
    const RecLog *x = (const RecLog *)a, *y = (const RecLog *)b;
    if (x->spec != y->spec) return x->spec < y->spec ? -1 : 1;
    if (x->host != y->host) return x->host < y->host ? -1 : 1;
    if (x->i != y->i) return x->i < y->i ? -1 : 1;
    return x->j < y->j ? -1 : (x->j > y->j);
}
ORC_EXPORT uint64_t orc_get_recombination_log(orc_world *w, uint32_t c, uint64_t *host, uint32_t *i, uint32_t *j, uint32_t *s0, uint32_t *s1, uint8_t *which) {
    Sim *s = &w->sims[c];
    if (!host) return s->n_rlog;
    qsort(s->rlog, s->n_rlog, sizeof(RecLog), cmp_rlog); /* processing order = (spec, host, i, j) */
    for (uint64_t k = 0; k < s->n_rlog; k++) { host[k] = s->rlog[k].host; i[k] = s->rlog[k].i; j[k] = s->rlog[k].j; s0[k] = s->rlog[k].s0; s1[k] = s->rlog[k].s1; which[k] = s->rlog[k].which; }
    return s->n_rlog;
}
ORC_EXPORT void orc_get_offspring(orc_world *w, uint32_t c, uint64_t *out) { memcpy(out, w->sims[c].offspring, w->sims[c].n * sizeof(uint64_t)); }
ORC_EXPORT void orc_get_replication_terms(orc_world *w, uint32_t c, double *fit, double *hsum, double *lam) {
    Sim *s = &w->sims[c];
    memcpy(fit, s->fit, s->n * 8); memcpy(hsum, s->hsum, s->n * 8); memcpy(lam, s->lam, s->n * 8);
}
static int find_provider(orc_world *w, uint32_t spec, int which) {
    if (spec >= w->n_specs) return -1;
    return which ? w->specs[spec].infect : w->specs[spec].repro;
}
ORC_EXPORT int orc_get_raw_fitness(orc_world *w, uint32_t c, uint32_t spec, int which, double *out) {
    Sim *s = &w->sims[c];
    int pi = find_provider(w, spec, which);
    for (uint64_t i = 0; i < s->n; i++) out[i] = pi >= 0 ? raw_fitness(w, pi, s->pop[i]) : 1.0;
    return 0;
}
ORC_EXPORT uint64_t orc_export_mutations(orc_world *w, uint32_t c, uint64_t *offsets, uint32_t *positions, uint8_t *bases) {
    Sim *s = &w->sims[c];
    MutSet M = {0};
    uint64_t total = 0;
    int64_t last_id = -1; /* consecutive duplicates are common */
    for (uint64_t i = 0; i < s->n; i++) {
        if (s->pop[i] != last_id) { hap_get_mutations(w, s->pop[i], &M); last_id = s->pop[i]; }
        if (offsets) offsets[i] = total;
        if (positions) for (uint32_t m = 0; m < M.n; m++) { positions[total + m] = M.m[m].pos; bases[total + m] = M.m[m].base; }
        total += M.n;
    }
    if (offsets) offsets[s->n] = total;
    free(M.m);
    return total;
}
ORC_EXPORT uint8_t orc_get_base(orc_world *w, uint32_t c, uint64_t infectant, uint32_t pos) { return hap_get_base(w, w->sims[c].pop[infectant], pos); }
ORC_EXPORT void orc_get_population_ids(orc_world *w, uint32_t c, int64_t *out) { memcpy(out, w->sims[c].pop, w->sims[c].n * sizeof(int64_t)); }

/* ----------------------------------- direct haplotype API (KATs, lib.rs:113-142) */
ORC_EXPORT int64_t orc_hap_wildtype(orc_world *w) { (void)w; return 0; }
ORC_EXPORT int64_t orc_hap_create_descendant(orc_world *w, int64_t parent, const uint32_t *pos, const uint8_t *to, uint32_t n) {
    int bad = 0;
    int64_t id = hap_create_descendant(w, parent, pos, to, n, &bad);
    return bad ? -1 - id : id; /* negative: assert_ne!(from,to) would have fired */
}
ORC_EXPORT int64_t orc_hap_create_recombinant(orc_world *w, int64_t l, int64_t r, uint32_t lp, uint32_t rp) { return hap_create_recombinant(w, l, r, lp, rp); }
ORC_EXPORT uint8_t orc_hap_get_base(orc_world *w, int64_t id, uint32_t pos) { return hap_get_base(w, id, pos); }
ORC_EXPORT void orc_hap_get_sequence(orc_world *w, int64_t id, uint8_t *out) { /* haplotype.rs:410-419 */
    MutSet M = {0};
    hap_get_mutations(w, id, &M);
    memcpy(out, w->wt, w->L);
    for (uint32_t m = 0; m < M.n; m++) out[M.m[m].pos] = M.m[m].base;
    free(M.m);
}
ORC_EXPORT uint32_t orc_hap_get_mutations(orc_world *w, int64_t id, uint32_t *pos, uint8_t *base, uint32_t cap) {
    MutSet M = {0};
    hap_get_mutations(w, id, &M);
    for (uint32_t m = 0; m < M.n && m < cap; m++) { pos[m] = M.m[m].pos; base[m] = M.m[m].base; }
    uint32_t n = M.n; free(M.m); return n;
}
ORC_EXPORT double orc_hap_raw_fitness(orc_world *w, int64_t id, uint32_t spec, int which) { int pi = find_provider(w, spec, which); return pi >= 0 ? raw_fitness(w, pi, id) : 1.0; }
ORC_EXPORT double orc_hap_fitness(orc_world *w, int64_t id, uint32_t spec, int which) { int pi = find_provider(w, spec, which); return pi >= 0 ? mapped_fitness(w, pi, id) : 1.0; }
ORC_EXPORT double orc_table_get_value(orc_world *w, uint32_t spec, int which, uint32_t pos, uint8_t base) { int pi = find_provider(w, spec, which); return tab(&w->prov[pi], pos, base); }
ORC_EXPORT double orc_utility_apply(const vl_utility *u, double f) { return utility_apply(u, f); }
/* put an arbitrary haplotype into a population slot (test set-up) */
ORC_EXPORT void orc_set_infectant(orc_world *w, uint32_t c, uint64_t i, int64_t id) { w->sims[c].pop[i] = id; }
ORC_EXPORT uint64_t orc_n_nodes(orc_world *w) { return w->n_nodes; }

/* samplers exposed for distribution tests of the restated rand_distr algorithms */
ORC_EXPORT void orc_test_poisson(uint64_t seed, double lambda, uint64_t n, double *out) { Rng r; rng_seed(&r, seed); int ok; for (uint64_t i = 0; i < n; i++) out[i] = rng_poisson(&r, lambda, &ok); }
ORC_EXPORT void orc_test_binomial(uint64_t seed, uint64_t nn, double p, uint64_t n, uint64_t *out) { Rng r; rng_seed(&r, seed); for (uint64_t i = 0; i < n; i++) out[i] = rng_binomial(&r, nn, p); }
ORC_EXPORT int orc_test_alias(uint64_t seed, const uint64_t *wts, uint64_t nw, uint64_t n, uint64_t *out) {
    Rng r; rng_seed(&r, seed); Alias a;
    if (alias_build(&a, wts, nw)) return -1;
    for (uint64_t i = 0; i < n; i++) out[i] = alias_sample(&a, &r);
    alias_free(&a); return 0;
}
ORC_EXPORT int orc_num_threads(orc_world *w) { return w->n_threads; }
ORC_EXPORT void orc_set_mode(orc_world *w, int mode) { w->mode = mode; }
