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


def test_replay_with_atomic_host_sums_within_tolerance():
    # VL_FLAG_ATOMIC_HOST_SUM accumulates per-host sums in arrival order: same values up to f64 rounding
    from virolution_b200.abi import VL_FLAG_ATOMIC_HOST_SUM
    c = helpers.case("crowded")
    w = helpers.make_oracle(c)
    gpu = helpers.make_gpu(c, flags=VL_FLAG_ATOMIC_HOST_SUM)
    for g in range(3):
        log = helpers.oracle_generation(w)
        gpu.increment_generation()
        gpu.replay_infections(log["infections"])
        gpu.infect()
        gpu.replay_mutations(log["mut_offsets"], log["mut_sites"], log["mut_bases"])
        gpu.mutate_infectants()
        gpu.replay_offspring(log["offspring"])
        gpu.replicate_infectants()
        f, s, l = gpu.get_replication_terms()
        helpers.assert_f64_equal(f, log["fit"], "fitness")
        helpers.assert_f64_equal(s, log["hsum"], "host sums", rtol=1e-12)
        helpers.assert_f64_equal(l, log["lam"], "lambda", rtol=1e-12)
        gpu.set_population(gpu.select(log["sample_idx"]))
    gpu.close()
