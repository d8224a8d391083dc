"""Replay parity: the CUDA path, fed the oracle's recorded draws through the C ABI, must reproduce every oracle
output bit for bit (integers equal, f64 bit patterns equal) — host map, mutation sets, raw fitness of every provider,
mapped fitness, per-host sums, lambda, offspring prefix, target size and the selected next population."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", helpers.ALL_CASES)
def test_replay_bitexact_against_live_oracle(name):
    c = helpers.case(name)
    w = helpers.make_oracle(c)
    gpu = helpers.make_gpu(c)
    n_specs = len(c["specs"])
    for g in range(c["generations"]):
        log = helpers.oracle_generation(w)
        helpers.replay_generation(gpu, log, n_specs)
    # the surviving population carries the same haplotypes in the same order
    eo, ep, eb = gpu.export_mutations()
    oo, op, ob = w.export_mutations()
    assert np.array_equal(eo, oo) and np.array_equal(ep, op) and np.array_equal(eb, ob)
    gpu.close()


@pytest.mark.parametrize("name", helpers.GOLDEN_CASES)
def test_replay_bitexact_against_golden_vectors(name):
    c = helpers.case(name)
    gold = dict(np.load(os.path.join(helpers.GOLDEN, f"oracle_replay_{name}.npz")))
    gpu = helpers.make_gpu(c)
    for g in range(c["generations"]):
        helpers.replay_generation(gpu, helpers.split_log(gold, g), len(c["specs"]))
    gpu.close()


def test_replay_survives_garbage_collection():
    c = helpers.case("epistasis_dense")
    w = helpers.make_oracle(c)
    gpu = helpers.make_gpu(c)
    for g in range(c["generations"]):
        helpers.replay_generation(gpu, helpers.oracle_generation(w), 1)
        before = gpu.get_counters()
        gpu.collect_garbage()
        after = gpu.get_counters()
        assert after["gc_runs"] == before["gc_runs"] + 1
        assert after["haplotype_records"] <= before["haplotype_records"]
        assert np.array_equal(gpu.get_raw_fitness().view(np.uint64), w.get_raw_fitness().view(np.uint64))
    gpu.close()


def test_hill_utility_within_tolerance():
    # Hill goes through pow(): CUDA's pow is not glibc's, so 1e-12 relative (north_star's fp tolerance) instead of bits
    c = helpers.case("hill")
    w = helpers.make_oracle(c)
    gpu = helpers.make_gpu(c)
    for g in range(c["generations"]):
        helpers.replay_generation(gpu, helpers.oracle_generation(w), 1, rtol=1e-12)
    gpu.close()


def test_replay_bitexact_at_1e5():
    """SURVEY.md Appendix C asks for a 10^5 run besides the K1-size cases: HIV-shaped parameters (hiv.yaml's substitution
    matrix shape, R0 6, dilution 1) on N = H = 10^5, three replayed generations, host tiles and multi-tile resampling."""
    from virolution_b200.abi import FitnessProvider, HostSpec, Parameters
    from virolution_b200.config import exponential_table
    wt = helpers.random_wildtype(2600, 31)
    n = 100_000
    par = Parameters(2e-4, 0.0, n, 6.0, n, 1.0, helpers.HIV_SUBST)
    tab = exponential_table(wt, rng=np.random.default_rng(32))
    c = dict(wildtype=wt, parameters=par, specs=[HostSpec(0, n, FitnessProvider("NonEpistatic", table=tab, utility=helpers.ALG))],
             n0=n, generations=3, seed=33)
    w = helpers.make_oracle(c)
    gpu = helpers.make_gpu(c)
    for g in range(c["generations"]):
        helpers.replay_generation(gpu, helpers.oracle_generation(w), 1)
    eo, ep, eb = gpu.export_mutations()
    oo, op, ob = w.export_mutations()
    assert np.array_equal(eo, oo) and np.array_equal(ep, op) and np.array_equal(eb, ob)
    gpu.close()
