"""Shared set-up for the parity tests: seeded cases, oracle recording and GPU replay/compare."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from virolution_b200.abi import FitnessProvider, HostSpec, Parameters, UtilityFunction  # noqa: E402
from virolution_b200.config import exponential_table  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_FIX = None

UNIFORM_SUBST = ((0, 1, 1, 1), (1, 0, 1, 1), (1, 1, 0, 1), (1, 1, 1, 0))
HIV_SUBST = ((0.0, 0.1, 0.2, 0.7), (0.1, 0.0, 0.6, 0.3), (0.2, 0.6, 0.0, 0.2), (0.6, 0.2, 0.2, 0.0))


def fixtures():
    global _FIX
    if _FIX is None:
        _FIX = dict(np.load(os.path.join(GOLDEN, "reference_fixtures.npz")))
    return _FIX


def random_wildtype(L, seed):
    return np.random.default_rng(seed).integers(0, 4, size=L).astype(np.uint8)


def synthetic_pairs(wt, n_pairs, seed):
    """(n,5) pair table: pos1 < pos2 uniform, non-wildtype bases, value ~ LogNormal(0, 0.1)."""
    rng = np.random.default_rng(seed)
    L = len(wt)
    rows = []
    for _ in range(n_pairs):
        p1, p2 = sorted(rng.choice(L, size=2, replace=False))
        b1 = (wt[p1] + rng.integers(1, 4)) % 4
        b2 = (wt[p2] + rng.integers(1, 4)) % 4
        rows.append((p1, b1, p2, b2, float(rng.lognormal(0.0, 0.1))))
    return np.array(rows, dtype=np.float64)


ALG = UtilityFunction("Algebraic", upper=1.5)


def case(name):
    """-> dict(wildtype, parameters, specs, n0, generations, seed)."""
    fx = fixtures()
    if name == "singlehost":  # tests/singlehost scale-down; mutation rate raised so every generation mutates
        wt = fx["phix174_wildtype"]
        par = Parameters(1e-4, 0.0, 3000, 100.0, 3000, 0.02, UNIFORM_SUBST)
        tab = exponential_table(wt, rng=np.random.default_rng(11))
        specs = [HostSpec(0, 3000, FitnessProvider("NonEpistatic", table=tab, utility=ALG))]
        return dict(wildtype=wt, parameters=par, specs=specs, n0=3000, generations=6, seed=1)
    if name == "crowded":  # many infectants per host (thread tier and heavy tier), lethal-heavy table, HIV matrix
        wt = random_wildtype(200, 5)
        par = Parameters(2e-3, 0.0, 60, 6.0, 4000, 1.0, HIV_SUBST)
        tab = exponential_table(wt, weights=(0.2, 0.3, 0.5, 0.0), rng=np.random.default_rng(12))
        specs = [HostSpec(0, 60, FitnessProvider("NonEpistatic", table=tab, utility=UtilityFunction("Linear")))]
        return dict(wildtype=wt, parameters=par, specs=specs, n0=4000, generations=6, seed=2)
    if name == "epistasis_fixture":  # tests/epistasis tables (2 pairs) on phiX174
        wt = fx["phix174_wildtype"]
        par = Parameters(2e-4, 0.0, 2000, 100.0, 2000, 0.02, UNIFORM_SUBST)
        specs = [HostSpec(0, 2000, FitnessProvider("SimpleEpistatic", table=fx["epistasis_fitness"],
                                                   epistasis=fx["epistasis_pairs"], utility=ALG))]
        return dict(wildtype=wt, parameters=par, specs=specs, n0=2000, generations=5, seed=3)
    if name == "epistasis_dense":  # short genome + many pairs so the pair rules fire constantly
        wt = random_wildtype(48, 6)
        par = Parameters(2e-2, 0.0, 800, 20.0, 1500, 0.1, UNIFORM_SUBST)
        tab = exponential_table(wt, weights=(0.4, 0.5, 0.0, 0.1), rng=np.random.default_rng(13))
        pairs = synthetic_pairs(wt, 400, 14)
        specs = [HostSpec(0, 800, FitnessProvider("SimpleEpistatic", table=tab, epistasis=pairs, utility=ALG))]
        return dict(wildtype=wt, parameters=par, specs=specs, n0=1500, generations=8, seed=4)
    if name == "multihost":  # two host specs with their own tables, infective fitness, n_hits > 1
        wt = random_wildtype(600, 7)
        par = Parameters(5e-4, 0.0, 1000, 30.0, 2500, 0.05, UNIFORM_SUBST, n_hits=3)
        # no lethal entries: a haplotype that is lethal in one spec but alive in the other can mutate the lethal site
        # again, and T[to]/T[from] with T[from] == 0 makes its raw fitness NaN — reference-undefined territory
        t0 = exponential_table(wt, weights=(0.3, 0.6, 0.0, 0.1), rng=np.random.default_rng(15))
        t1 = exponential_table(wt, weights=(0.3, 0.6, 0.0, 0.1), rng=np.random.default_rng(16))
        ti = exponential_table(wt, weights=(0.1, 0.8, 0.0, 0.1), rng=np.random.default_rng(17))
        for t in (t0, t1, ti):  # (deleterious entries are max(1 - Exp, 0): about one in a hundred comes out exactly 0)
            np.maximum(t, 0.02, out=t)
        specs = [HostSpec(0, 500, FitnessProvider("NonEpistatic", table=t0, utility=ALG),
                          FitnessProvider("NonEpistatic", table=ti, utility=UtilityFunction("Algebraic", upper=1.2))),
                 HostSpec(500, 1000, FitnessProvider("NonEpistatic", table=t1, utility=UtilityFunction("Linear")), None)]
        return dict(wildtype=wt, parameters=par, specs=specs, n0=2500, generations=6, seed=5)
    if name == "recombination":  # rho > 0, several infectants per host
        wt = random_wildtype(300, 8)
        par = Parameters(1e-3, 2e-2, 150, 8.0, 1500, 0.5, UNIFORM_SUBST)
        tab = exponential_table(wt, weights=(0.3, 0.6, 0.0, 0.1), rng=np.random.default_rng(18))
        pairs = synthetic_pairs(wt, 300, 19)
        specs = [HostSpec(0, 150, FitnessProvider("SimpleEpistatic", table=tab, epistasis=pairs, utility=ALG))]
        return dict(wildtype=wt, parameters=par, specs=specs, n0=1500, generations=6, seed=6)
    if name == "neutral":  # data/hiv.yaml shape: Neutral + Linear, dilution 1, R0 6
        wt = fx["hiv_wildtype"]
        par = Parameters(2.16e-5, 0.0, 500, 6.0, 500, 1.0, HIV_SUBST)
        specs = [HostSpec(0, 500, FitnessProvider("Neutral"))]
        return dict(wildtype=wt, parameters=par, specs=specs, n0=500, generations=5, seed=7)
    if name == "hill":  # Hill utility goes through pow(): tolerance instead of bit equality
        wt = random_wildtype(400, 9)
        par = Parameters(1e-3, 0.0, 1000, 10.0, 1000, 0.2, UNIFORM_SUBST)
        tab = exponential_table(wt, rng=np.random.default_rng(20))
        specs = [HostSpec(0, 1000, FitnessProvider("NonEpistatic", table=tab, utility=UtilityFunction("Hill", p=0.5, n=2.0)))]
        return dict(wildtype=wt, parameters=par, specs=specs, n0=1000, generations=4, seed=8)
    raise KeyError(name)


ALL_CASES = ["singlehost", "crowded", "epistasis_fixture", "epistasis_dense", "multihost", "recombination", "neutral"]
GOLDEN_CASES = ["crowded", "epistasis_dense", "recombination"]


def make_oracle(c, n_compartments=1, n_threads=1):
    from oracle.oracle import OracleWorld
    w = OracleWorld(c["wildtype"], c["parameters"], c["specs"], n_compartments=n_compartments, seed=c["seed"], n_threads=n_threads)
    for k in range(n_compartments):
        w.set_population_wildtype(c["n0"], k)
    return w


def make_gpu(c, seed=None, flags=0, compartment_id=0, device=0, **caps):
    from virolution_b200.simulation import GpuSimulation
    g = GpuSimulation(c["wildtype"], c["parameters"], c["specs"], seed=c["seed"] if seed is None else seed, flags=flags,
                      compartment_id=compartment_id, device=device, **caps)
    g.set_population_wildtype(c["n0"])
    return g


def oracle_generation(w, comp=0):
    """One generation on the oracle, returning every recorded draw and every checked output."""
    log = {}
    w.increment_generation(comp)
    n = w.population_len(comp)
    log["n"] = np.array([n], dtype=np.uint64)
    w.infect(comp)
    log["infections"] = w.get_infections(comp)
    off, hosts = w.get_host_map(comp)
    log["hostmap_offsets"], log["hostmap_hosts"] = off, hosts
    w.mutate_infectants(comp)
    mo, ms, mb = w.get_mutation_log(comp)
    log["mut_offsets"], log["mut_sites"], log["mut_bases"] = mo, ms, mb
    rh, ri, rj, rs0, rs1, rw = w.get_recombination_log(comp)
    log["rec_host"], log["rec_i"], log["rec_j"], log["rec_s0"], log["rec_s1"], log["rec_which"] = rh, ri, rj, rs0, rs1, rw
    offspring = w.replicate_infectants(comp)
    log["offspring"] = offspring
    f, s, l = w.get_replication_terms(comp)
    log["fit"], log["hsum"], log["lam"] = f, s, l
    n_specs = w._cfg.cfg.n_specs
    for sp in range(n_specs):
        for which in (0, 1):
            log[f"raw_{sp}_{which}"] = w.get_raw_fitness(sp, which, comp)
    eo, ep, eb = w.export_mutations(comp)
    log["exp_offsets"], log["exp_pos"], log["exp_base"] = eo, ep, eb
    log["target_size"] = np.array([w.target_size(offspring, comp)])
    pop, idx = w.subsample_population(offspring, 1.0, comp)
    log["sample_idx"] = idx
    w.set_population([pop], comp)
    log["n_after"] = np.array([w.population_len(comp)], dtype=np.uint64)
    return log


def record_oracle_run(name):
    c = case(name)
    w = make_oracle(c)
    out = {}
    for g in range(c["generations"]):
        for k, v in oracle_generation(w).items():
            out[f"g{g}_{k}"] = v
    return out


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_f64_equal(a, b, what, rtol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{what}: NaN pattern differs"
    if rtol == 0.0:
        bad = np.nonzero((bits(a) != bits(b)) & ~nan_a)[0]
        assert bad.size == 0, f"{what}: {bad.size} values differ bitwise, first at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}"
    else:
        ok = np.isclose(a[~nan_a], b[~nan_a], rtol=rtol, atol=0.0)
        assert ok.all(), f"{what}: relative error above {rtol}"


def replay_generation(gpu, log, n_specs, rtol=0.0, check_prefix=True):
    """Feed one recorded generation into the CUDA path through the C ABI and compare every output with the log."""
    n = int(log["n"][0])
    assert gpu.population_len() == n
    gpu.increment_generation()
    gpu.replay_infections(log["infections"])
    gpu.infect()
    assert np.array_equal(gpu.get_infections(), log["infections"])
    off, hosts = gpu.get_host_map()
    assert np.array_equal(off, log["hostmap_offsets"]), "host map offsets"
    assert np.array_equal(hosts, log["hostmap_hosts"]), "host map order (descending infectant index inside a host)"
    gpu.replay_recombinations(log["rec_host"], log["rec_i"], log["rec_j"], log["rec_s0"], log["rec_s1"], log["rec_which"])
    gpu.replay_mutations(log["mut_offsets"], log["mut_sites"], log["mut_bases"])
    gpu.mutate_infectants()
    eo, ep, eb = gpu.export_mutations()
    assert np.array_equal(eo, log["exp_offsets"]), "mutation set sizes"
    assert np.array_equal(ep, log["exp_pos"]) and np.array_equal(eb, log["exp_base"]), "mutation sets"
    for sp in range(n_specs):
        for which in (0, 1):
            assert_f64_equal(gpu.get_raw_fitness(sp, which), log[f"raw_{sp}_{which}"], f"raw fitness spec {sp} which {which}", rtol)
    gpu.replay_offspring(log["offspring"])
    offspring = gpu.replicate_infectants()
    assert np.array_equal(offspring, log["offspring"])
    f, s, l = gpu.get_replication_terms()
    assert_f64_equal(f, log["fit"], "mapped fitness", rtol)
    assert_f64_equal(s, log["hsum"], "per-host fitness sum", rtol)
    # the device keeps the mean it actually draws from: 0.0 wherever Poisson::new fails (lambda not finite or <= 0)
    lam_o = log["lam"]
    with np.errstate(invalid="ignore"):
        lam_expect = np.where(np.isnan(log["fit"]), np.nan, np.where(np.isfinite(lam_o) & (lam_o > 0), lam_o, 0.0))
    assert_f64_equal(l, lam_expect, "lambda", rtol)
    if check_prefix:
        assert np.array_equal(gpu.get_offspring_prefix(), np.cumsum(log["offspring"], dtype=np.uint64)), "offspring prefix"
    ts = gpu.target_size()
    assert ts == float(log["target_size"][0]), "target size"
    assert gpu.target_size(log["offspring"]) == float(log["target_size"][0]), "target size from a host offspring map"
    gpu.set_population(gpu.select(log["sample_idx"]))
    assert gpu.population_len() == int(log["n_after"][0])


def split_log(npz, g):
    p = f"g{g}_"
    return {k[len(p):]: v for k, v in npz.items() if k.startswith(p)}


def mutation_sets(offsets, pos, base):
    """list of tuples of (pos, base) per infectant."""
    out = []
    for i in range(len(offsets) - 1):
        a, b = int(offsets[i]), int(offsets[i + 1])
        out.append(tuple(zip(pos[a:b].tolist(), base[a:b].tolist())))
    return out
