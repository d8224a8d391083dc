"""The device-resident loop the bench times (vl_run: tile-strided generation, vl_tiles.cuh) against the trait-level
calls and against the general fused kernels, at sizes where every fast kernel is live (N = H >= 3*10^5: bucketed
host map, integer-threshold mutation counts, fused replicate, host tiles) and at the BASELINE config shapes.
Equality is exact: same seed => same population, infectant by infectant, same raw fitness bit for bit."""
import numpy as np
import pytest

import helpers
from virolution_b200 import abi
from virolution_b200.workloads import workload

pytestmark = pytest.mark.gpu


def make(wl, seed, flags=0):
    from virolution_b200.simulation import GpuSimulation
    g = GpuSimulation(wl["wildtype"], wl["parameters"], wl["specs"], seed=seed, flags=flags)
    g.set_population_wildtype(wl["n0"])
    return g


def state(g):
    off, pos, base = g.export_mutations()
    return off, pos, base, g.get_raw_fitness().view(np.uint64)


def assert_same_population(a, b, what):
    sa, sb = state(a), state(b)
    assert a.population_len() == b.population_len(), what
    for x, y, name in zip(sa, sb, ("mutation offsets", "positions", "bases", "raw fitness bits")):
        assert np.array_equal(x, y), f"{what}: {name} differ"


def stepwise(g, k):
    for _ in range(k):  # Runner::run's loop body through the trait-level calls (src/runner.rs:382-422), offspring kept on the device
        g.increment_generation()
        g.infect()
        g.mutate_infectants()
        g.replicate_infectants(to_host=False)
        g.set_population(g.subsample_population(None, 1.0))


@pytest.mark.parametrize("name,n,gens", [("k2_hiv", 300_000, 6), ("k1_singlehost", 300_000, 5), ("k4_epistasis", 300_000, 4)])
def test_vl_run_equals_trait_level_calls_at_scale(name, n, gens):
    wl = workload(name, n)
    if name == "k1_singlehost":  # the shipped mutation rate gives ~5 mutants per generation at this size: raise it so records matter
        wl["parameters"].mutation_rate = 2e-5
    a, b, c = make(wl, 77), make(wl, 77), make(wl, 77, flags=abi.VL_FLAG_NO_TILES)
    a.run(gens)
    stepwise(b, gens)
    c.run(gens)
    assert a.get_generation() == gens
    assert a.infectant_generations() == b.infectant_generations() == c.infectant_generations()
    assert_same_population(a, b, f"{name}: vl_run (tile-strided generation) vs trait-level calls")
    assert_same_population(a, c, f"{name}: tile-strided generation vs general fused kernels")
    off = state(a)[0]
    assert (np.diff(off.astype(np.int64)) > 0).sum() > 1000, "the run must have produced mutants"
    # continue both fused arms from the same state: the carried per-infectant fitness must keep agreeing with the memo
    a.run(3)
    c.run(3)
    assert_same_population(a, c, f"{name}: three more generations")
    for g in (a, b, c):
        g.close()


def test_vl_run_after_trait_level_generations_and_back():
    """Switching between the loops mid-run: the per-infectant fitness array is rebuilt from the haplotype memo whenever
    the population was installed by anything but the tile-strided resample."""
    wl = workload("k2_hiv", 200_000)
    a, b = make(wl, 5), make(wl, 5, flags=abi.VL_FLAG_NO_TILES)
    for g in (a, b):
        g.run(2)
        stepwise(g, 1)
        g.run(2)
        g.collect_garbage()
        g.run(2)
    assert_same_population(a, b, "mixed loops")
    a.close()
    b.close()


def test_vl_run_full_k2_properties():
    """K2 at 10^6 (BASELINE configs[1]): size-independent properties of the timed loop."""
    wl = workload("k2_hiv", 1_000_000)
    g = make(wl, 11)
    n0 = wl["n0"]
    g.run(8)
    assert g.population_len() == n0          # R0 = 6, dilution 1: the cap max_population binds every generation
    assert g.infectant_generations() == 8 * n0
    off, pos, base = g.export_mutations()
    assert np.all(base != wl["wildtype"][pos]), "mutation sets never hold the wildtype base"
    owner = np.repeat(np.arange(n0), np.diff(off.astype(np.int64)))
    same = owner[1:] == owner[:-1]
    assert np.all(np.diff(pos.astype(np.int64))[same] > 0), "positions ascending inside every infectant"
    # mean number of mutations per infectant after g generations ~ g * mu * L, less what selection removed; bounded above
    mean_mut = pos.size / n0
    assert 0.5 < mean_mut < 8 * 2.16e-5 * 9700 * 1.2
    raw = g.get_raw_fitness()
    assert np.all(np.isfinite(raw)) and np.all(raw >= 0)
    g.close()


def test_events_of_many_sites_take_the_second_pass():
    """mu * L = 5: one event in fifteen hits more than 8 sites — more than the per-thread shared scratch of the main k_mutate
    pass holds.  The device-resident generation sets those aside for a second pass (same record ids, same draws); the
    general fused kernels and the trait-level calls handle them in place.  Same populations."""
    from test_gpu_stochastic import big_singlehost
    c = big_singlehost(n=120_000, mu=5.0 / 3000, seed=3)
    wl = dict(wildtype=c["wildtype"], parameters=c["parameters"], specs=c["specs"], n0=c["n0"])
    a, b, d = make(wl, 19), make(wl, 19), make(wl, 19, flags=abi.VL_FLAG_NO_TILES)
    a.run(4)
    stepwise(b, 4)
    d.run(4)
    assert_same_population(a, b, "many-site events: vl_run vs trait-level calls")
    assert_same_population(a, d, "many-site events: device-resident vs general fused kernels")
    off = state(a)[0]
    assert np.diff(off.astype(np.int64)).max() > 12
    for g in (a, b, d):
        g.close()
